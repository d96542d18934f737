"""Time hole_rank on the config-3 shape for both precisions (select the library with HOLE_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine, HOLE_RANK_BF16, HOLE_RANK_BF16X3, HOLE_SIDE_TAIL

kg = D.make_config("diffbot_d256", n_triples=200000)
e = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E)
Q = int(os.environ.get("QUERIES", 100000))
q = torch.from_numpy(kg.triples[:Q]).cuda()
for name, prec in (("bf16", HOLE_RANK_BF16), ("bf16x3", HOLE_RANK_BF16X3)):
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            e.rank(q, HOLE_SIDE_TAIL, kg.n_relations, kg.n_rows, precision=prec)
        except Exception as ex:
            print(name, "unsupported:", str(ex)[:80]); best = None; break
        e1.record(); torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    if best:
        flops = 2.0 * 256 * Q * kg.n_entities * (3 if prec == HOLE_RANK_BF16X3 else 1)
        print(f"{name}: {best:.2f} ms  {Q * kg.n_entities / best / 1e9:.2f} e12 scores/s  {flops / best / 1e9:.0f} TFLOP/s issued")
