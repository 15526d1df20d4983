"""Concurrent pinned host <-> device bandwidth on every GPU of the box at once (the denominator of bench.py's e2e numbers at N > 1):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 scripts/h2d_d2h_bench.py
Every rank copies 1 GiB pinned -> device and 1 GiB device -> pinned, first one direction at a time, then both at once on two
streams, all ranks simultaneously (barrier before each phase).  Prints per-rank and aggregate GB/s."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 27   # 1 GiB of float64
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.ones(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def phase(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return n * 8 / (ms * 1e-3) / 1e9   # GB/s per direction


res = [phase(True, False), phase(False, True), phase(True, True)]
t = torch.tensor(res, dtype=torch.float64, device=dev)
if world > 1:
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    names = ["H2D alone", "D2H alone", "H2D + D2H at once (per direction)"]
    print("%d GPUs, 1 GiB pinned transfers, all ranks simultaneously" % world)
    for i, nm in enumerate(names):
        vals = [float(a[i]) for a in allr]
        print("  %-36s per rank min %.1f / max %.1f GB/s, aggregate %.1f GB/s" % (nm, min(vals), max(vals), sum(vals)))
if world > 1:
    dist.destroy_process_group()
