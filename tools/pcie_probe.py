"""PCIe ceiling of the box for the e2e leg: pinned H2D, D2H and both at once (same sizes as cfg2: 1.47 GB in, 0.74 GB out)."""
import time
import torch

torch.cuda.init()
h_in = torch.empty(1474560000 // 4, dtype=torch.float32).pin_memory()
h_out = torch.empty(737114112 // 4, dtype=torch.float32).pin_memory()
d_in = torch.empty_like(h_in, device="cuda")
d_out = torch.empty_like(h_out, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {a*1e3:.2f} ms ({h_in.numel()*4/a/1e9:.1f} GB/s)  D2H {b*1e3:.2f} ms ({h_out.numel()*4/b/1e9:.1f} GB/s)  both {c*1e3:.2f} ms")
