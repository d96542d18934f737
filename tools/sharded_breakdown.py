"""Where does a row-sharded training step spend its time?  (torchrun, exploration)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from graphembeddings_b200 import data as D
from graphembeddings_b200 import sharded as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Bl, steps = 32768, 12
kg = D.make_config("diffbot_d256", n_triples=Bl * world * steps)
off, ids = D.build_type_csr(kg.type_of)
be = S.CudaBackend(kg.n_relations, kg.dim, Bl, local, kg.type_of, off, ids)
tr = (S.RowShardedTrainer if os.environ.get("HOLE_SHARDED_NCCL") == "1" else S.P2PRowShardedTrainer)(kg.n_relations, kg.n_entities, kg.dim, be, dist).load_embeddings(kg.E)
tri = torch.from_numpy(kg.triples).view(steps, world, Bl, 3)[:, rank].contiguous().cuda()
S.TIMING = {}
for s in range(steps):
    if s == 4:
        S.TIMING = {}
    tr.train_step(tri[s], 1, s, 0.2, 0.1)
torch.cuda.synchronize()
if rank == 0:
    tot = sum(S.TIMING.values())
    for k, v in S.TIMING.items():
        print(f"{k:28s} {v / (steps - 4) * 1e3:8.1f} us/step  {v / tot:6.1%}")
    print(f"{'total':28s} {tot / (steps - 4) * 1e3:8.1f} us/step")
dist.barrier(); dist.destroy_process_group()
