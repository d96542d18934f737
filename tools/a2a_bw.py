"""all_to_all_single bandwidth at the sharded step's message size (torchrun)."""
import os, sys, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = 98304
x = torch.randn(rows, 256, device="cuda"); y = torch.empty_like(x)
per = rows // world
for _ in range(5):
    dist.all_to_all_single(y, x, [per] * world, [per] * world)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_to_all_single(y, x, [per] * world, [per] * world)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
remote = x.numel() * 4 * (world - 1) / world
if rank == 0:
    print(f"NCHAN={os.environ.get('NCCL_MAX_P2P_NCHANNELS','default')} a2a {x.numel()*4/1e6:.0f} MB/rank: {ms*1e3:.0f} us, remote bytes {remote/1e6:.0f} MB -> {remote/ms/1e6:.0f} GB/s per direction")
dist.destroy_process_group()
