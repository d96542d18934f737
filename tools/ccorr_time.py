import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine
B=32768
kg = D.make_config("diffbot_d256", n_triples=13*B, trained_scale=True)
off, ids = D.build_type_csr(kg.type_of)
eng = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
eng.set_relation_count(kg.n_relations)
eng.set_score_mode("ccorr_tanh")
v, ms = bench.timed_train(eng, torch.from_numpy(kg.triples).cuda(), B, 10, 3, 16, first_step=2000)
print("ccorr_tanh d=256 B=32768: %.1f M triples/s, %.3f ms/step, %.1f TFLOP/s fp32" % (v/1e6, ms, v*32*128*128/1e12))
