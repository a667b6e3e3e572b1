"""Does a stream that waits (eagerly, after graph.replay()) on an EXTERNAL event recorded inside the graph wait for
THIS replay's record node?  Kernel A (long) -> record ev -> kernel B (long) in the graph; the side stream waits for ev and
stamps a flag copy; we check that the side stream's work ran after A of the same replay and before B finished."""
import torch
dev = torch.device('cuda')
x = torch.zeros(1 << 28, device=dev)          # 1 GB: each pass ~0.3 ms
flag = torch.zeros(1, device=dev)
seen = torch.zeros(8, device=dev)
ev = torch.cuda.Event(external=True)
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(4): x.add_(1.0)            # A
    flag.add_(1.0)
    ev.record()
    for _ in range(8): x.add_(1.0)            # B
    flag.add_(100.0)
torch.cuda.synchronize()
flag.zero_()
e_side, e_end, e_start = (torch.cuda.Event(enable_timing=True) for _ in range(3))
for it in range(4):
    e_start.record()
    g.replay()
    with torch.cuda.stream(side):
        side.wait_event(ev)
        seen[it].copy_(flag[0])               # expect (it * 101 + 1): A of THIS replay done, B not
        e_side.record()
    e_end.record()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    print(f'replay {it}: side saw flag {seen[it].item():.0f} (want {it * 101 + 1}); side done at {e_start.elapsed_time(e_side):.2f} ms, graph done at {e_start.elapsed_time(e_end):.2f} ms')
