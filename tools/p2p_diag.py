"""Diagnostic: which way of mapping a peer rank's buffer lets our kernels write to it?
torchrun --nproc-per-node 2 tools/p2p_diag.py   (P2P_VARIANT=A: tensor on the peer device, B: opened on mine)"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.multiprocessing.reductions import reduce_tensor
from graphembeddings_b200.engine import HoleEngine

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
variant = os.environ.get("P2P_VARIANT", "A")
eng = HoleEngine(4096, 64, lr)
W = torch.full((4096, eng.row_stride), float(rank + 1), device=f"cuda:{lr}")
everyone = [None] * dist.get_world_size()
dist.all_gather_object(everyone, (reduce_tensor(W), lr))
peer = 1 - rank
(fn, args), dev_k = everyone[peer]
args = list(args)
print(rank, "variant", variant, "storage_device in handle:", args[6], flush=True)
if variant == "B":
    args[6] = lr
eng.enable_peer_access(dev_k)
P = fn(*args)
print(rank, "peer tensor device", P.device, "ptr", hex(P.data_ptr()), flush=True)
torch.cuda.synchronize(); dist.barrier()
print(rank, "torch read of peer:", float(P[:4].sum().item()), flush=True)
src = torch.full((100, eng.row_stride), 10.0 + rank, device=f"cuda:{lr}")
ids = torch.arange(100, device=f"cuda:{lr}", dtype=torch.int64)
eng.gather_rows(src, ids, 0, P[1000:1100])
torch.cuda.synchronize(); dist.barrier()
print(rank, "my W after peer's push:", float(W[1000:1100].mean().item()), "(expect", 10.0 + peer, ")", flush=True)
eng.add_rows(src, ids, 0, P[0:100])
torch.cuda.synchronize(); dist.barrier()
print(rank, "pull+add:", float(src.mean().item()), "(expect", 10.0 + rank + peer + 1, ")", flush=True)
dist.destroy_process_group()
