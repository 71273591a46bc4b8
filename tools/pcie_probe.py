"""PCIe probe: pinned H2D alone, D2H alone, both directions at once (what bounds bench.py's e2e)."""
import json
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


def chunks(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1):
                d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)
    return f


res = {"gib": n / 2**30}
res["h2d_gbs"] = n / timed(h2d) / 1e6
res["d2h_gbs"] = n / timed(d2h) / 1e6
t = timed(both)
res["both_each_gbs"] = n / t / 1e6
res["both_16chunks_each_gbs"] = n / timed(chunks(16)) / 1e6
print(json.dumps(res))
