"""torchrun check on N GPUs: row-sharded training and sharded ranking against a single-GPU
engine run on rank 0 (same global batches).  Prints OK lines; exits non-zero on mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine
from graphembeddings_b200.sharded import CudaBackend, RowShardedTrainer, P2PRowShardedTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Bl, steps = 2048, int(os.environ.get("STEPS", 4))
kg = D.synthetic_kg(9, 50000, Bl * world * steps, 5, 256, seed=77, trained_scale=True)
off, ids = D.build_type_csr(kg.type_of)
be = CudaBackend(kg.n_relations, kg.dim, Bl, local, kg.type_of, off, ids)
from graphembeddings_b200.sharded import make_trainer
tr = make_trainer(kg.n_relations, kg.n_entities, kg.dim, be, dist, log=print).load_embeddings(kg.E)
if rank == 0:
    print("trainer:", type(tr).__name__)
ahead = os.environ.get("NO_AHEAD") != "1"        # work one step ahead on the side stream (host tensors in)


def slice_of(s):
    return torch.from_numpy(kg.triples[(s * world + rank) * Bl:(s * world + rank + 1) * Bl])


for s in range(steps):
    tr.train_step(slice_of(s), 3, s, 0.2, 0.1, next_pos=slice_of(s + 1) if (ahead and s + 1 < steps) else None)
full = tr.gather_embeddings()
q = torch.from_numpy(kg.triples[:1000])
raw, filt = tr.rank(q, 0)
ok = True
if rank == 0:
    e = HoleEngine(kg.n_rows, kg.dim, local).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
    for s in range(steps):
        gb = kg.triples[s * Bl * world:(s + 1) * Bl * world]
        side, neg = e.corrupt_batch(gb, 3, s)
        e.train_step(gb, neg, side, 0.2, 0.1)
    want = e.embeddings()
    err = float((full - want).abs().max())
    print(f"train: max |sharded - single| = {err:.3e} (moved {float((want.cpu() - torch.from_numpy(kg.E)).abs().max()):.3e})")
    ok &= err < 2e-6
    # ranking on the single-GPU table that equals the gathered sharded table up to ~1e-7
    e2 = HoleEngine(kg.n_rows, kg.dim, local).set_embeddings(full)
    r1, f1, _ = e2.rank(q, 0, kg.n_relations, kg.n_rows)
    same = float((r1 == raw).float().mean())
    print(f"rank: identical counts for {same:.4f} of queries; max |d| = {int((r1 - raw).abs().max())}")
    ok &= same > 0.999
    print("MULTI_GPU_CHECK", "OK" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
