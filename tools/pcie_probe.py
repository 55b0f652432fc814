'''Host <-> device copy bandwidth of the box (pinned memory, 67 MB = the state of a 128^3 cavity): one direction at a
time and both at once -- the ceiling of the `e2e` leg of bench.py.  python tools/pcie_probe.py'''
import torch

n = 128 ** 3 * 4
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device='cuda')
d_out = torch.zeros(n, dtype=torch.float64, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


mb = n * 8 / 1e6
for name, fn in (('H2D', h2d), ('D2H', d2h), ('both at once', both)):
    ms = timed(fn)
    print('%-13s %.3f ms per %.0f MB (each way) = %.1f GB/s per direction' % (name, ms, mb, mb / ms))
