"""Slab-grid software pipelining of the Z and Y stages (ACCFFT_B200_OVERLAP = chunks): timing and bitwise comparison
with the unpipelined path, one process per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/overlap_bench.py [n=512]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import accfft_b200 as acc

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
shape = (n, n, n)
ref_out = None
for chunks in [1, 2, 4, 8, 16, 1]:
    os.environ["ACCFFT_B200_OVERLAP"] = str(chunks)
    comm = acc.accfft_create_comm([world, 1])
    p = acc.Plan(shape, comm, r2c=True, single=False)
    isz = p.isize
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    x = torch.rand(isz[0] * isz[1] * isz[2], dtype=torch.float64, device=dev, generator=gen)
    out = torch.zeros(p.alloc_max // 8, dtype=torch.float64, device=dev)
    back = torch.zeros_like(x)
    st = torch.cuda.Stream()
    p.set_stream(st.cuda_stream); p.set_blocking(False)
    def step():
        p.execute_r2c(x, out); p.execute_c2r(out, back)
    with torch.cuda.stream(st):
        for _ in range(5): step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if ref_out is None:
        ref_out = out.clone()
    same = bool(torch.equal(out, ref_out))
    err = float(torch.linalg.vector_norm(back / n ** 3 - x) / torch.linalg.vector_norm(x))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 40
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(iters): step()
        e1.record(st)
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = torch.tensor([int(same and err < 1e-12)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        N = n ** 3
        print("overlap chunks=%2d grid %dx1: %.4f ms/round-trip  %.0f GFLOP/s  spectrum bitwise equal to chunks=1 and round trip ok on all ranks: %s" %
              (chunks, world, float(ms.item()), 2 * 5 * N * np.log2(N) / (float(ms.item()) * 1e-3) / 1e9, bool(ok.item())), flush=True)
    p.destroy()
dist.destroy_process_group()
