import sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench
r = bench.bottleneck_leg(torch.device('cuda', 0))
print({k: round(v, 1) if isinstance(v, float) else v for k, v in r['large'].items() if k in ('fwd_us', 'bwd_us', 'fwd_gbs', 'bwd_gbs')})
