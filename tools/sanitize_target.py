"""Small end-to-end workload for compute-sanitizer (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine, HOLE_SIDE_BOTH

for dim, B in ((150, 777), (256, 1024), (64, 300)):
    kg = D.synthetic_kg(7, 1500, B * 3, 4, dim, seed=dim, trained_scale=True, zipf_entities=True)
    off, ids = D.build_type_csr(kg.type_of)
    e = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
    e.set_relation_count(kg.n_relations)
    side, neg = e.corrupt_batch(kg.triples[:B], 3, 0)
    e.train_step(kg.triples[:B], neg, side, 0.2, 0.1)
    e.train_steps(kg.triples, B, 3, 1, 0.2, [0.1] * 3)
    e.train_steps_host(kg.triples, B, 3, 4, 0.2, [0.1] * 3)
    e.evaluate_triples(kg.triples[:100])
    raw, filt, ts = e.rank(kg.triples[:300], HOLE_SIDE_BOTH, kg.n_relations, kg.n_rows)
    torch.cuda.synchronize()
    e.close()
print("sanitize target ok")
