"""Why is the first 20-step call of a context slower than the following ones?  Prints us/step of a sequence of
hole_train_steps calls of different lengths on BASELINE config 1 (B = 32768), with and without idle gaps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B_
from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine

B = 32768
kg = D.make_config(B_.WORKLOAD, n_triples=400 * B)
off, ids = D.build_type_csr(kg.type_of)
eng = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
eng.set_relation_count(kg.n_relations)
tri = torch.from_numpy(kg.triples).cuda()
B_.settle_clocks()
SEQ = eval(os.environ.get('PROBE_SEQ', '[(5, 0), (20, 0), (20, 0), (5, 0), (20, 0), (20, 0), (60, 0), (20, 0), (20, 0.2), (20, 0)]'))
k0 = 0
out = []
for n, gap in SEQ:
    lrs = B_.lr_schedule(n, k0, 30_000_000 // B)
    if gap:
        time.sleep(gap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.train_steps(tri[k0 * B:(k0 + n) * B], B, 1, k0, B_.MARGIN, lrs)
    e1.record()
    torch.cuda.synchronize()
    out.append((n, gap, round(e0.elapsed_time(e1) * 1e3 / n, 2)))
    k0 += n
print(out)
