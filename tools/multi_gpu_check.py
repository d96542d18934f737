"""torchrun check on N GPUs (real CUDA-IPC peers over NVLink): row-sharded training
(P2PRowShardedTrainer = hole_shard_step) and sharded ranking against a single-GPU engine run on
rank 0 over the same global batches.  Prints one JSON line on rank 0; exits non-zero on mismatch.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from graphembeddings_b200.sharded import parity_case


def run_case(rank, world, local, dim, Bl, steps, n_ent, seed, chunked):
    ok, res = parity_case(dist, rank, world, local, dim, Bl, steps, n_ent, seed, chunked, log=print)
    if rank == 0:
        print("MULTI_GPU_CHECK " + json.dumps(res), flush=True)
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [(256, 2048, 4, 50000, 77, False), (150, 777, 5, 3000, 78, True), (64, 300, 3, 500, 79, False)]
    if os.environ.get("CHECK_QUICK") == "1":
        cases = cases[:2]
    for rep in range(int(os.environ.get("CHECK_REPEAT", 1))):
        for dim, Bl, steps, n_ent, seed, chunked in cases:
            ok &= run_case(rank, world, local, dim, Bl, steps, n_ent, seed + 100 * rep, chunked)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
