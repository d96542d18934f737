"""A/B: sharded step with and without the fast path (delta kernels / side-stream plan), async timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from graphembeddings_b200 import data as D
from graphembeddings_b200 import sharded as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Bl, steps = 32768, 24
kg = D.make_config("diffbot_d256", n_triples=Bl * world * steps)
off, ids = D.build_type_csr(kg.type_of)
be = S.CudaBackend(kg.n_relations, kg.dim, Bl, local, kg.type_of, off, ids)
tr = S.RowShardedTrainer(kg.n_relations, kg.n_entities, kg.dim, be, dist).load_embeddings(kg.E)
tri = torch.from_numpy(kg.triples).view(steps, world, Bl, 3)[:, rank].contiguous().cuda()

def run(tag, noplan=False):
    if noplan:
        be.plan = lambda a, b: None
    for s in range(4):
        tr.train_step(tri[s], 1, s, 0.2, 0.1)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(4, steps):
        tr.train_step(tri[s], 1, s, 0.2, 0.1)
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / (steps - 4)
    if rank == 0:
        print(f"{tag:32s} {dt * 1e6:9.1f} us/step  {Bl * world / dt / 1e6:8.1f} M triples/s", flush=True)

run("NCCL all-to-all path")
tr2 = S.P2PRowShardedTrainer(kg.n_relations, kg.n_entities, kg.dim, be, dist).load_embeddings(kg.E)
tr_nccl, tr = tr, tr2
run("P2P push/pull path")
tr = tr_nccl
run("fast path, plan inline", noplan=True)
del S.CudaBackend.step_delta
run("generic path")
dist.barrier(); dist.destroy_process_group()
