"""NCCL all-to-all bandwidth at the exchange sizes of the 512^3 double transform (measurement tooling):
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/micro/nccl_a2a.py"""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for group_size in (2, 4, 8):
    if group_size > world:
        continue
    groups = [dist.new_group(list(range(b, b + group_size))) for b in range(0, world, group_size)]
    grp = groups[rank // group_size]
    for mb in (134.7,):   # bytes of one rank's whole local array at 2x4
        n = int(mb * 1e6 / 16 / group_size) * group_size
        src = torch.zeros(n, 2, dtype=torch.float64, device="cuda")
        dst = torch.zeros_like(src)
        for _ in range(3):
            dist.all_to_all_single(dst, src, group=grp)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_to_all_single(dst, src, group=grp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out = n * 16 * (group_size - 1) / group_size
        if rank == 0:
            print(f"nccl all_to_all group {group_size}: {n*16/1e6:.1f} MB/rank, {ms:.3f} ms, NVLink out {out/ms/1e6:.1f} GB/s per GPU", flush=True)
dist.destroy_process_group()
