"""GPU-timeline breakdown of the sharded step WITHOUT host synchronisation: stream events at
the section boundaries + the host's enqueue time per step.
torchrun --nproc-per-node N tools/sharded_timeline.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from graphembeddings_b200 import data as D
from graphembeddings_b200 import sharded as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Bl, steps, warm = int(os.environ.get("BATCH", 32768)), 24, 4
kg = D.make_config("diffbot_d256", n_triples=Bl * world * steps)
off, ids = D.build_type_csr(kg.type_of)
be = S.CudaBackend(kg.n_relations, kg.dim, Bl, local, kg.type_of, off, ids)
cls = S.RowShardedTrainer if os.environ.get("HOLE_SHARDED_NCCL") == "1" else S.P2PRowShardedTrainer
tr = cls(kg.n_relations, kg.n_entities, kg.dim, be, dist).load_embeddings(kg.E)
tri = torch.from_numpy(kg.triples).view(steps, world, Bl, 3)[:, rank].contiguous().cuda()
for s in range(warm):
    tr.train_step(tri[s], 1, s, 0.2, 0.1)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
S.EVENTS = []
marks = []
t0 = time.perf_counter()
for s in range(warm, steps):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    tr.train_step(tri[s], 1, s, 0.2, 0.1)
e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
host = (time.perf_counter() - t0) / (steps - warm)
torch.cuda.synchronize()
n = steps - warm
tot = marks[0].elapsed_time(marks[-1]) / n * 1e3
acc = {}
for name, a, b in S.EVENTS:
    acc[name] = acc.get(name, 0.0) + a.elapsed_time(b) * 1e3 / n
if rank == 0:
    print(f"{cls.__name__}: world {world}, batch/rank {Bl}")
    for k, v in acc.items():
        print(f"  {k:40s} {v:8.1f} us")
    print(f"  {'(between sections)':40s} {tot - sum(acc.values()):8.1f} us")
    print(f"  GPU step {tot:.1f} us; host enqueue {host * 1e6:.1f} us/step; {Bl * world / tot:.1f} M triples/s")
dist.barrier(); dist.destroy_process_group()
