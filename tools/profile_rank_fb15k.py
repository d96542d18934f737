import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine, HOLE_SIDE_BOTH
kg = D.make_config("rank_fb15k_d150", trained_scale=True)
e = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E)
for _ in range(2):
    raw, filt, ts = e.rank(kg.triples, HOLE_SIDE_BOTH, kg.n_relations, kg.n_rows)
torch.cuda.synchronize()
print("ok", int(raw.max()))
