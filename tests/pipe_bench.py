"""Per-SM pipe throughput micro-benchmark (debug entry point svae_debug_pipe_bench): cycles per warp instruction."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sparse_vae_b200 import _native as N  # noqa: E402

DBG = N.load_debug()      # libsvae_b200_dbg.so: `python sparse_vae_b200/csrc/build.py --debug`

out = torch.zeros(64, dtype=torch.int64, device='cuda')
iters = 256
names = ['MUFU.EX2', 'F2FP bf16x2 pack', 'FFMA', 'FMNMX3', 'tcgen05.ld x32 (4 KB)', 'tcgen05.st x16 (2 KB)', 'softmax step (4 elem)', 'softmax step, 96 elem unrolled', 'fwd softmax pass 2 (160 scores in regs)', 'fwd softmax pass 1 + 2 (5 LDTM + max + exp)']
print(f'cycles per warp instruction (modes 6, 7: per PAIR of elements; modes 8, 9: cycles per block-row softmax), {iters} x 8 instructions per warp')
for mode, name in enumerate(names):
    row = []
    for warps in (1, 4, 8):
        for _ in range(2):
            out.zero_()
            N.check(DBG.svae_debug_pipe_bench(mode, warps, iters, out.data_ptr(), torch.cuda.current_stream().cuda_stream), 'bench')
            torch.cuda.synchronize()
        cyc = out[:warps].max().item() / (iters * {6: 4, 7: 48, 8: 1, 9: 1}.get(mode, 8))
        row.append(f'{warps:2d} warps {cyc:7.2f}')
    print(f'  {name:24s} ' + ' | '.join(row))
print('(4 warps = one per SMSP; 8 / 16 warps = 2 / 4 per SMSP sharing that SMSP\'s pipes and TMEM lane quarter)')

print('with a 5th / 9th warp issuing back-to-back S-shaped tcgen05.mma (M=128, N=256, K=64) into the TMEM columns being read:')
for mode, name in ((4, 'tcgen05.ld x32 (4 KB)'), (5, 'tcgen05.st x16 (2 KB)'), (8, 'fwd softmax pass 2'), (9, 'fwd softmax pass 1 + 2')):
    row = []
    for warps in (5,):
        for _ in range(2):
            out.zero_()
            N.check(DBG.svae_debug_pipe_bench(mode | 0x100, warps, iters, out.data_ptr(), torch.cuda.current_stream().cuda_stream), 'bench')
            torch.cuda.synchronize()
        cyc = out[:warps - 1].max().item() / (iters * {6: 4, 7: 48, 8: 1, 9: 1}.get(mode, 8))
        row.append(f'{warps - 1} warps + MMA {cyc:8.2f}  (MMA groups issued meanwhile: {out[62].item()})')
    print(f'  {name:24s} ' + ' | '.join(row))
