cd /root/repo
for v in "" _bn3 _bn5 _bn6; do echo "variant '$v'"; SVAE_LIB_VARIANT=$v python tools/dbg/bottleneck_leg.py 2>&1 | tail -1; done
