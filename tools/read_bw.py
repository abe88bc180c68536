import torch, time
x = torch.ones(1<<30, dtype=torch.float64, device="cuda")  # 8 GB
for fn, name, nbytes in ((lambda: x.sum(), "sum fp64 (read only)", x.numel()*8), (lambda: x.view(torch.int64).max(), "max int64 (read only)", x.numel()*8), (lambda: x.add_(1.0), "add_ in place (read+write)", 2*x.numel()*8)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(f"{name}: {ms:.3f} ms -> {nbytes/ms/1e6:.0f} GB/s")
