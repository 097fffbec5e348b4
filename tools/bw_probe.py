"""HBM write-only / copy bandwidth probe with torch ops (context for the write-bound gas-optics roofline)."""
import torch
n = 2 * 1024**3  # floats -> 8 GiB
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_()); print(f"memset  {4*n/ms/1e6:8.1f} GB/s (write only)")
ms = t(lambda: x.fill_(1.5)); print(f"fill    {4*n/ms/1e6:8.1f} GB/s (write only)")
ms = t(lambda: y.copy_(x)); print(f"copy    {8*n/ms/1e6:8.1f} GB/s (read + write)")
ms = t(lambda: x.sum()); print(f"sum     {4*n/ms/1e6:8.1f} GB/s (read only)")
