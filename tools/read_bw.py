"""Read-only streaming bandwidth of this GPU (context for the cell-stats kernel, which only reads its 12 B/pixel):
torch reductions over a 944 MB float32 tensor (the size of one 256-frame VGA batch) and a 4 GB one, plus a copy."""
import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mb in (944, 4096):
    x = torch.empty(mb * 250000, dtype=torch.float32, device="cuda").normal_()
    y = torch.empty_like(x)
    ms = t(lambda: x.sum())
    print(f"{mb} MB read (sum):  {x.numel()*4/ms/1e6:8.1f} GB/s")
    ms = t(lambda: x.max())
    print(f"{mb} MB read (max):  {x.numel()*4/ms/1e6:8.1f} GB/s")
    ms = t(lambda: y.copy_(x))
    print(f"{mb} MB copy (r+w):  {2*x.numel()*4/ms/1e6:8.1f} GB/s")
    ms = t(lambda: y.zero_())
    print(f"{mb} MB write (zero): {x.numel()*4/ms/1e6:8.1f} GB/s")
