"""Write-only HBM bandwidth on this GPU (the denominator that bounds mesh_emit, which writes 10x
what it reads): cudaMemset through torch, 2 GiB, best of 10."""
import torch
x = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
best = {}
for name, fn in (("memset (zero_)", lambda: x.zero_()), ("fill_ int32", lambda: x.view(torch.int32).fill_(7)),
                 ("copy (read+write)", lambda: y.copy_(x))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    b = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        b = min(b, e0.elapsed_time(e1))
    nbytes = x.numel() * (2 if "copy" in name else 1)
    print("%-20s %.3f ms  %.0f GB/s" % (name, b, nbytes / b / 1e6))
