"""torchrun check on N GPUs (real CUDA-IPC peers over NVLink): row-sharded training
(P2PRowShardedTrainer = hole_shard_step) and sharded ranking against a single-GPU engine run on
rank 0 over the same global batches.  Prints one JSON line on rank 0; exits non-zero on mismatch.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine
from graphembeddings_b200.sharded import CudaBackend, P2PRowShardedTrainer, make_trainer


def run_case(rank, world, local, dim, Bl, steps, n_ent, seed, chunked):
    kg = D.synthetic_kg(9, n_ent, Bl * world * steps, 5, dim, seed=seed, trained_scale=True, zipf_entities=True)
    off, ids = D.build_type_csr(kg.type_of)
    be = CudaBackend(kg.n_relations, kg.dim, Bl, local, kg.type_of, off, ids)
    tr = make_trainer(kg.n_relations, kg.n_entities, kg.dim, be, dist, log=print).load_embeddings(kg.E)
    mine = torch.from_numpy(kg.triples).view(steps, world, Bl, 3)[:, rank].contiguous().cuda()
    lrs = [0.1 / (1 + 0.01 * s) for s in range(steps)]
    if chunked and isinstance(tr, P2PRowShardedTrainer):
        tr.train_steps(mine.view(-1, 3), Bl, 3, 0, 0.2, lrs)
    else:
        for s in range(steps):
            tr.train_step(mine[s], 3, s, 0.2, lrs[s], next_pos=mine[s + 1] if s + 1 < steps else None)
    full = tr.gather_embeddings()
    q = torch.from_numpy(kg.triples[:1000])
    raw, filt = tr.rank(q, 0)
    res = {"trainer": type(tr).__name__, "world": world, "dim": dim, "batch_per_gpu": Bl, "steps": steps,
           "chunked_call": bool(chunked)}
    ok = True
    if rank == 0:
        e = HoleEngine(kg.n_rows, kg.dim, local).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
        for s in range(steps):
            gb = kg.triples[s * Bl * world:(s + 1) * Bl * world]
            side, neg = e.corrupt_batch(gb, 3, s)
            e.train_step(gb, neg, side, 0.2, lrs[s])
        want = e.embeddings()
        err = float((full - want).abs().max())
        moved = float((want.cpu() - torch.from_numpy(kg.E)).abs().max())
        # ranking on the single-GPU table that equals the gathered sharded table up to ~1e-7
        e2 = HoleEngine(kg.n_rows, kg.dim, local).set_embeddings(full)
        r1, f1, _ = e2.rank(q, 0, kg.n_relations, kg.n_rows)
        same = float((r1 == raw).float().mean())
        res.update({"max_abs_err": err, "moved": moved, "rank_counts_equal": same,
                    "rank_counts_max_diff": int((r1 - raw).abs().max())})
        ok = err <= 2e-6 and moved > 1e-4 and same == 1.0
        res["ok"] = ok
        print("MULTI_GPU_CHECK " + json.dumps(res), flush=True)
        e.close()
        e2.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [(256, 2048, 4, 50000, 77, False), (150, 777, 5, 3000, 78, True), (64, 300, 3, 500, 79, False)]
    if os.environ.get("CHECK_QUICK") == "1":
        cases = cases[:2]
    for dim, Bl, steps, n_ent, seed, chunked in cases:
        ok &= run_case(rank, world, local, dim, Bl, steps, n_ent, seed, chunked)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
