cd /root/repo; mkdir -p gpurun_out
python tools/dbg/test_norm_probe.py 2>&1 | grep "^\[" 
python -m pytest tools/dbg/test_norm_probe.py -q -s -m "" 2>&1 | grep "^\["
python -m pytest tools/dbg/test_norm_probe.py -q -s -p no:hypothesispytest -p no:cacheprovider 2>&1 | grep "^\["
