"""What the box's PCIe link gives pinned-memory copies (the ceiling of bench.py's host-buffer e2e leg):
H2D alone, D2H alone, and both directions at once on two streams.   python profiles/pcie_ceiling.py"""
import json

import torch

n = 1 << 28                                  # 1 GiB per buffer
h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


gb = 4 * n / 1e9
print(json.dumps({"h2d_gbs": round(gb / timed(h2d), 1), "d2h_gbs": round(gb / timed(d2h), 1),
                  "bidirectional_each_gbs": round(gb / timed(both), 1),
                  "note": "bench.py e2e moves 11.1 GB each way per step: ms_per_step >= 11.1 / bidirectional_each_gbs"}))
