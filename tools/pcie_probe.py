#!/usr/bin/env python
"""PCIe ceiling of the box: pinned H2D, D2H and both at once (GB/s).  The e2e leg of bench.py moves
8 MiB per block in each direction, so its bound is min(H2D, D2H under concurrent load) / 8 MiB."""
import json
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_in, do_out, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if do_in:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(True, True, 1)
res = {"h2d_GBps": run(True, False), "d2h_GBps": run(False, True), "both_each_GBps": run(True, True)}
res["e2e_blocks_per_s_bound_n4096"] = res["both_each_GBps"] * 1e9 / (8 << 20)
print(json.dumps(res))
