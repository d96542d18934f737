"""What HBM delivers for K1's ACCESS PATTERN: random 1 KB rows of a 1.2 GB table read (and written back),
with the library's plain row kernels (no arithmetic), against a streaming copy of the same volume.
Prints GB/s of DRAM-side bytes (each row counted once per read and once per write)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from graphembeddings_b200.engine import HoleEngine

N, D = 1_200_014, 256
eng = HoleEngine(N, D)
eng.table = torch.randn((N, eng.row_stride), device="cuda")
gen = torch.Generator(device="cuda")
gen.manual_seed(0)
out = {}
for n in (98304, 393216):
    ids = torch.randperm(N, device="cuda", generator=gen)[:n].to(torch.int64)        # distinct random rows
    dst = torch.empty((n, eng.row_stride), device="cuda")
    rows = torch.randn((n, eng.row_stride), device="cuda")
    row_b = eng.row_stride * 4

    def timed(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best * 1e-3

    t = timed(lambda: eng.gather_rows(eng.table, ids, 0, dst))
    out[f"gather_{n}"] = {"us": t * 1e6, "GBps": 2 * n * row_b / t / 1e9, "pattern": "random 1 KB row read + sequential write"}
    t = timed(lambda: eng.add_rows(eng.table, ids, 0, rows))
    out[f"add_{n}"] = {"us": t * 1e6, "GBps": 3 * n * row_b / t / 1e9,
                       "pattern": "random 1 KB row read + sequential row read + random 1 KB row write (K1's in-place update)"}
    a = torch.empty((3 * n, eng.row_stride), device="cuda")
    b = torch.empty_like(a)
    t = timed(lambda: b.copy_(a))
    out[f"copy_{n}"] = {"us": t * 1e6, "GBps": 2 * a.numel() * 4 / t / 1e9, "pattern": "streaming copy of 3n rows"}
print(json.dumps(out, indent=1))
