import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / 'tests'))
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

def probe(tag):
    import test_gpu_data_parallel_nccl as T
    dev = torch.device('cuda', 0)
    model = T._model(dev)
    for it in range(2):
        loss = T._step(model, T._batch(dev, 0, T.B_GLOBAL)).item()
        gs = {n: p.grad.norm().item() for n, p in model.named_parameters() if p.grad is not None}
        tot = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).norm().item()
        print(f'[{tag}] it {it} loss {loss!r} gnorm {tot!r} tf32 {torch.backends.cuda.matmul.allow_tf32} prec {torch.get_float32_matmul_precision()} threads {torch.get_num_threads()}')
    top = sorted(gs.items(), key=lambda kv: -kv[1])[:6]
    print(f'[{tag}]', ' '.join(f'{n}={v:.6e}' for n, v in top))

def test_probe():
    probe('pytest')

if __name__ == '__main__':
    probe('plain')
