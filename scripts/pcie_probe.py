"""Pinned host<->device copy rates at the e2e transfer sizes (is the e2e arm at the PCIe limit?)."""
import torch

dev = torch.device("cuda:0")
for mb_in, mb_out in ((2.13, 1.75), (8.5, 7.0), (34.0, 28.0)):
    n_in, n_out = int(mb_in * 1e6 / 4), int(mb_out * 1e6 / 4)
    h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
    d_out = torch.empty(n_out, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def t(fn, reps=50):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d(); d2h()

    for f in (h2d, d2h, both):
        f()
    a, b, c = t(h2d), t(d2h), t(both)
    print(f"H2D {mb_in} MB: {a:.1f} us ({mb_in * 1e3 / a:.1f} GB/s) | D2H {mb_out} MB: {b:.1f} us ({mb_out * 1e3 / b:.1f} GB/s) | both concurrently: {c:.1f} us")
