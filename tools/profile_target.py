"""Short workload for ncu: a few training steps of BASELINE config 1 at the bench batch size,
then one 20k-query ranking call on the same 1.2M-row table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine

B, steps = 32768, int(os.environ.get("PT_STEPS", 6))
kg = D.make_config("diffbot_d256", n_triples=B * steps)
off, ids = D.build_type_csr(kg.type_of)
e = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
e.set_relation_count(kg.n_relations)
tri = torch.from_numpy(kg.triples).cuda()
# two calls: the plan of the first one reports how many duplicated uses a step has, the second one then
# takes the one-launch sort (as every call after the first does in a training run)
e.train_steps(tri[:2 * B], B, 1, 0, 0.2, [0.1] * 2)
torch.cuda.synchronize()
sums = e.train_steps(tri[2 * B:], B, 1, 2, 0.2, [0.1] * (steps - 2))
torch.cuda.synchronize()
q = kg.triples[:20000]
raw, filt, ts = e.rank(q, 0, kg.n_relations, kg.n_rows)
torch.cuda.synchronize()
print("ok", float(sums.mean()) / B, int(raw.max()))
