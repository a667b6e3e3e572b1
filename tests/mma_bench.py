"""tcgen05.mma issue-cost micro-benchmark (debug entry point svae_debug_mma_bench)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sparse_vae_b200 import _native as N  # noqa: E402

DBG = N.load_debug()      # libsvae_b200_dbg.so: `python sparse_vae_b200/csrc/build.py --debug`

out = torch.zeros(4, dtype=torch.int64, device='cuda')
count = 64
print(f'{count} back-to-back MMAs, M=128 K=16: cycles per MMA (issue | issue+drain)')
for variant, name in ((0, 'SS K-major B'), (2, 'SS MN-major B'), (1, 'TS K-major B'), (3, 'TS MN-major B'),
                      (4, 'SS, two warps'), (7, 'TS MN, two warps'), (8, 'SS conv'), (9, 'TS conv'),
                      (11, 'TS MN conv'), (12, 'SS conv two warps'), (15, 'TS MN conv two'),
                      (16, 'SS unrolled'), (18, 'SS MN unrolled'), (19, 'TS MN unrolled'), (20, 'SS unrolled two'),
                      (23, 'TS MN unrolled two')):
    for n in (32, 64, 96, 128, 256):
        for _ in range(2):
            out.zero_()
            N.check(DBG.svae_debug_mma_bench(variant, n, count, out.data_ptr(), torch.cuda.current_stream().cuda_stream), 'bench')
            torch.cuda.synchronize()
        o = out.tolist()
        print(f'  {name:18s} N={n:3d}: warp0 {o[0] / count:7.1f} | {o[1] / count:7.1f}   warp1 {o[2] / count:7.1f} | {o[3] / count:7.1f}')
