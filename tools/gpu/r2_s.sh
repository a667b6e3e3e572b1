set -x
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_attention.py tests/test_gpu_bottleneck.py -m gpu -x -q 2>&1 | tail -3
for v in "" _novec _vec2; do
SVAE_LIB_VARIANT=$v python - <<PY
import torch, bench, json
print('variant "$v"', json.dumps({k: ({kk: round(vv, 1) if isinstance(vv, float) else vv for kk, vv in x.items() if kk != 'note'} if isinstance(x, dict) else x) for k, x in bench.bottleneck_leg(torch.device('cuda')).items()}))
PY
done
