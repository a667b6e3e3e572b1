cd /root/repo
echo pipelined; python tools/dbg/rotary_pair_bench.py 2>&1 | tail -3
echo plain; SVAE_LIB_VARIANT=_np python tools/dbg/rotary_pair_bench.py 2>&1 | tail -3
