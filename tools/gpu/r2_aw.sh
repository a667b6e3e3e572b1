cd /root/repo; python tools/dbg/sample_time.py 2>&1 | tail -5
cd /root/repo/_old; python ../tools/dbg/sample_time.py 2>&1 | tail -5
