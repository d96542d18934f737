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
Bl, steps, warm = int(os.environ.get("BATCH", 32768)), int(os.environ.get("STEPS", 24)), 4
kg = D.make_config("diffbot_d256", n_triples=Bl * world * steps)
off, ids = D.build_type_csr(kg.type_of)
be = S.CudaBackend(kg.n_relations, kg.dim, Bl, local, kg.type_of, off, ids)
cls = S.RowShardedTrainer if os.environ.get("HOLE_SHARDED_NCCL") == "1" else S.P2PRowShardedTrainer
tr = cls(kg.n_relations, kg.n_entities, kg.dim, be, dist).load_embeddings(kg.E)
tri = torch.from_numpy(kg.triples).view(steps, world, Bl, 3)[:, rank].contiguous().cuda()
VAR = os.environ.get("VARIANT", "").split(",")
if "pin" in VAR:
    host_tri = tri.cpu().pin_memory()
sampler = None
if "sampler" in VAR and rank == 0:
    import bench as B_
    sampler = B_.ClockSampler(local, period=0.25)
    sampler.start()
    time.sleep(1.0)
if "lr" in VAR:
    import bench as B_
    lrs = B_.lr_schedule(3 * steps, 0, 30_000_000 // (Bl * world))
    LR = lambda s: float(lrs[s])
else:
    LR = lambda s: 0.1
for s in range(warm):
    tr.train_step(tri[s], 1, s, 0.2, LR(s), next_pos=tri[s + 1] if "warmahead" in VAR else None)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
S.EVENTS = [] if os.environ.get('NO_EVENTS') != '1' else None
if os.environ.get('PROFILE') == '1':
    be.eng.profile(True)
marks = []
prof = None
if os.environ.get("CPROFILE") == "1" and rank == 0:
    import cProfile
    prof = cProfile.Profile()
    prof.enable()
t0 = time.perf_counter()
for s in range(warm, steps):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    tr.train_step(tri[s], 1, s, 0.2, LR(s), next_pos=tri[s + 1] if (s + 1 < steps and os.environ.get('NO_AHEAD') != '1') else None)
e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
host = (time.perf_counter() - t0) / (steps - warm)
if prof is not None:
    prof.disable()
    import pstats, io
    buf = io.StringIO()
    pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(14)
    print(buf.getvalue())
torch.cuda.synchronize()
n = steps - warm
if os.environ.get('PROFILE') == '1' and rank == 0:
    k1, k3, ns = be.eng.profile_read()
    print(f"  library events: K1 {k1 / ns * 1e3:.1f} us, K3 {k3 / ns * 1e3:.1f} us per step ({ns} steps)")
tot = marks[0].elapsed_time(marks[-1]) / n * 1e3
acc = {}
for name, a, b in (S.EVENTS or []):
    acc[name] = acc.get(name, 0.0) + a.elapsed_time(b) * 1e3 / n
if rank == 0:
    print(f"{cls.__name__}: world {world}, batch/rank {Bl}, variant {VAR}")
    for k, v in acc.items():
        print(f"  {k:40s} {v:8.1f} us")
    print(f"  {'(between sections)':40s} {tot - sum(acc.values()):8.1f} us")
    print(f"  GPU step {tot:.1f} us; host enqueue {host * 1e6:.1f} us/step; {Bl * world / tot:.1f} M triples/s")
dist.barrier(); dist.destroy_process_group()
