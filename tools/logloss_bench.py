"""Throughput of the --log_loss step (hole_train_step_logloss) on BASELINE config 1's table.
python tools/logloss_bench.py [B] [k] [l2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = 40
kg = D.make_config("diffbot_d256", n_triples=B * steps)
off, ids = D.build_type_csr(kg.type_of)
e = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
tri = torch.from_numpy(kg.triples).to(e.device)
for k, l2 in ((1, 0.0), (2, 0.0), (4, 0.0), (1, 1e-9)):
    for s in range(4):
        e.train_step_logloss(tri[s * B:(s + 1) * B], 1, s, 0.1, l2, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(4, steps):
        e.train_step_logloss(tri[s * B:(s + 1) * B], 1, s, 0.1, l2, k)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (steps - 4) * 1e3
    print(f"log-loss step B={B} k={k} l2={l2:g}: {us:8.1f} us/step  {B / us:7.1f} M positives/s  "
          f"{B * (1 + k) / us:7.1f} M loss rows/s", flush=True)
