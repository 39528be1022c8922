"""Host-to-device copy rate from pinned memory on this box, by copy size (CUDA events, median of 20)."""
import torch
dev = torch.device("cuda:0")
for mb in (0.64, 1.28, 2.56, 5.12, 20.48, 163.84):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    h.normal_()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{mb:8.2f} MB  {ts[10] * 1000:8.1f} us  {mb / ts[10]:6.2f} GB/s")
