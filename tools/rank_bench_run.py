import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphembeddings_b200 import rank_bench, data as D
from graphembeddings_b200.engine import HoleEngine
full = "--full" in sys.argv
eng = kg = None
if full:
    kg = D.make_config("diffbot_d256", n_triples=200000)
    eng = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E)
print(json.dumps(rank_bench.run(eng, kg, quick=False), indent=1))
