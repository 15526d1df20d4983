"""Multi-GPU parity + timing check, one process per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu/mgpu_check.py
Compares every rank's block with the numpy oracle (rank 0 generates the global field from a seed; every rank
regenerates it locally -- no data collective is needed for the check)."""
import os, sys, json
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
import accfft_b200 as acc
from oracle import accfft_oracle as orc

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
GRIDS = {2: [(2, 1), (1, 2)], 4: [(2, 2), (1, 4), (4, 1)], 8: [(2, 4), (4, 2), (8, 1)]}[world]
worst = 0.0


def check(n, dims, r2c, single=False):
    global worst
    comm = acc.accfft_create_comm(list(dims))
    p = acc.Plan(n, comm, r2c=r2c, single=single)
    rdt, cdt = (np.float32, np.complex64) if single else (np.float64, np.complex128)
    x = orc.synthetic_real(n, 5) if r2c else orc.synthetic_complex(n, 5)
    X = orc.forward_r2c(x) if r2c else orc.forward_c2c(x)
    loc = orc.local_input(x, n, dims, comm.coords).astype(rdt if r2c else cdt)
    xd = torch.from_numpy(np.ascontiguousarray(loc)).to(dev) if loc.size else torch.zeros(2, dtype=torch.float64, device=dev)
    es = 4 if single else 8
    out = torch.zeros(max(p.alloc_max // es, 2), dtype=torch.float32 if single else torch.float64, device=dev)
    back = torch.zeros_like(xd)
    errs = []
    for it in range(2):
        if r2c:
            p.execute_r2c(xd, out); p.execute_c2r(out, back)
        else:
            p.execute_c2c(acc.ACCFFT_FORWARD, xd, out); p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)
    osz = p.osize
    cnt = osz[0] * osz[1] * osz[2]
    want = orc.local_output(X, n, dims, comm.coords, r2c=r2c)
    if cnt:
        got = out.cpu().numpy().view(cdt)[:cnt].reshape(osz)
        errs.append(orc.rel_l2(got, want))
    if loc.size:
        errs.append(orc.rel_l2(back.cpu().numpy().reshape(loc.shape) / x.size, loc))
    if r2c and loc.size:
        la = torch.zeros_like(xd)
        p.laplace(la, xd)
        errs.append(orc.rel_l2(la.cpu().numpy().reshape(loc.shape), orc.local_input(orc.laplace(x), n, dims, comm.coords)))
        g = [torch.zeros_like(xd) for _ in range(3)]
        p.grad(g[0], g[1], g[2], xd)
        gs = [torch.zeros_like(xd) for _ in range(3)]
        p.grad_slow(gs[0], gs[1], gs[2], xd)
        dv, dvs = torch.zeros_like(xd), torch.zeros_like(xd)
        p.divergence(dv, g[0], g[1], g[2])
        p.divergence_slow(dvs, g[0], g[1], g[2])
        # the per-axis scheme drops the mode pair +-floor(N/2), the full-3-D ("slow") scheme only X == N/2 on the complex axes:
        # different operators for odd N (pinned by the reference-run fixtures, tests/golden/ref_*.npz) -- each has its oracle
        gx = [orc.grad(x, ax) for ax in range(3)]
        gxs = [orc.grad_slow(x, ax) for ax in range(3)]
        want_div = orc.local_input(orc.divergence(*gx), n, dims, comm.coords)
        want_divs = orc.local_input(orc.divergence_slow(*gx), n, dims, comm.coords)
        for ax in range(3):
            errs.append(orc.rel_l2(g[ax].cpu().numpy().reshape(loc.shape), orc.local_input(gx[ax], n, dims, comm.coords)))
            errs.append(orc.rel_l2(gs[ax].cpu().numpy().reshape(loc.shape), orc.local_input(gxs[ax], n, dims, comm.coords)))
        errs.append(orc.rel_l2(dv.cpu().numpy().reshape(loc.shape), want_div))
        errs.append(orc.rel_l2(dvs.cpu().numpy().reshape(loc.shape), want_divs))
    elif r2c:   # empty rank: still takes part in every collective call
        la = torch.zeros_like(xd); p.laplace(la, xd)
        g = [torch.zeros_like(xd) for _ in range(3)]; p.grad(g[0], g[1], g[2], xd); p.grad_slow(g[0], g[1], g[2], xd)
        p.divergence(la, g[0], g[1], g[2]); p.divergence_slow(la, g[0], g[1], g[2])
    if r2c:
        # single-axis real transforms (accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t): per-rank half spectra + round trip
        for axis in (0, 1, 2):
            p.execute_r2c_axis(axis, xd, out)
            ref = orc.axis_local_spectrum(orc.axis_forward(x, axis), n, dims, comm.coords, axis)
            b2 = torch.zeros_like(xd)
            p.execute_c2r_axis(axis, out, b2)
            if ref.size:
                errs.append(orc.rel_l2(out.cpu().numpy().view(cdt)[:ref.size].reshape(ref.shape), ref))
            if loc.size:
                errs.append(orc.rel_l2(b2.cpu().numpy().reshape(loc.shape) / n[axis], loc))
    e = max(errs) if errs else 0.0
    t = torch.tensor([e], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = max(worst, float(t.item()) / (1e7 if single else 1.0))
    if rank == 0:
        print(f"n={n} dims={dims} r2c={r2c} single={single}: max rel-L2 over ranks {float(t.item()):.2e}", flush=True)
    p.destroy()


QUICK = os.environ.get("MGPU_QUICK", "0") == "1"   # 8-GPU visits are charged 8x: fewer grids / sizes, same code paths
for dims in (GRIDS[:1] + GRIDS[-1:] if QUICK else GRIDS):
    for n in ([(16, 12, 20), (5, 6, 6), (128, 96, 160)] if QUICK else [(16, 12, 20), (64, 64, 64), (48, 40, 36), (5, 6, 6), (128, 96, 160)]):
        check(n, dims, True)
        if not QUICK or n[0] == 16:
            check(n, dims, False)
    check((64, 64, 64), dims, True, single=True)


def bench(n, dims, r2c=True, iters=20, single=False):
    comm = acc.accfft_create_comm(list(dims))
    p = acc.Plan(n, comm, r2c=r2c, single=single)
    isz = p.isize
    cnt = isz[0] * isz[1] * isz[2] * (1 if r2c else 2)
    rdt = torch.float32 if single else torch.float64
    x = torch.rand(cnt, dtype=rdt, device=dev)
    out = torch.zeros(p.alloc_max // (4 if single else 8), dtype=rdt, device=dev)
    back = torch.zeros_like(x)
    st = torch.cuda.Stream()
    p.set_stream(st.cuda_stream); p.set_blocking(False); p.set_profiling(True)
    def step():
        if r2c:
            p.execute_r2c(x, out); p.execute_c2r(out, back)
        else:
            p.execute_c2c(acc.ACCFFT_FORWARD, x, out); p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)
    with torch.cuda.stream(st):
        for _ in range(3): step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); p.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(iters): step()
        e1.record(st)
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    prof = p.profile_read()
    N = n[0] * n[1] * n[2]
    gf = 2 * 5 * N * np.log2(N) / (float(ms.item()) * 1e-3) / 1e9
    if rank == 0:
        print(f"bench n={n} dims={dims} r2c={r2c} single={single}: {float(ms.item()):.3f} ms/round-trip  {gf:.0f} GFLOP/s  stages(ms/launch): " +
              " ".join(f"{k}={v[0]/max(v[1],1):.3f}" for k, v in prof.items()), flush=True)
    if rank == 0:
        print("   stage kernels:", p.stage_kernels(), flush=True)
    if r2c and os.environ.get("MGPU_OPS", "1") == "1" and dims in ((2, 4), (8, 1), (2, 1), (2, 2)):
        # operator timings (report as time, not GFLOP/s)
        la = torch.zeros_like(x)
        g = [torch.zeros_like(x) for _ in range(3)]
        for name, fn in (("laplace", lambda: p.laplace(la, x)), ("grad (per-axis)", lambda: p.grad(g[0], g[1], g[2], x)),
                         ("grad_slow (3-D)", lambda: p.grad_slow(g[0], g[1], g[2], x)),
                         ("grad x", lambda: p.grad(g[0], None, None, x, xyz=1)), ("grad y", lambda: p.grad(None, g[1], None, x, xyz=2)),
                         ("grad z", lambda: p.grad(None, None, g[2], x, xyz=4)),
                         ("divergence (per-axis)", lambda: p.divergence(la, g[0], g[1], g[2])),
                         ("divergence_slow (3-D)", lambda: p.divergence_slow(la, g[0], g[1], g[2]))):
            with torch.cuda.stream(st):
                fn()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            with torch.cuda.stream(st):
                e0.record(st)
                for _ in range(5): fn()
                e1.record(st)
            torch.cuda.synchronize(); dist.barrier()
            ms = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"   op {name}: {float(ms.item()):.3f} ms", flush=True)
    p.destroy()


for dims in GRIDS + ([(1, 8)] if world == 8 else []):
    bench((512, 512, 512), dims, True)
for dims in (GRIDS[:1] + GRIDS[-1:] if world == 8 else GRIDS[:2]):
    bench((512, 512, 512), dims, False)          # BASELINE config 3: 512^3 double c2c
if world == 8:
    bench((1024, 1024, 1024), (2, 4), True, single=True)   # config 4 (+ operators)
    bench((1024, 1024, 1024), (8, 1), True, single=True)
    bench((768, 768, 384), (2, 4), True)                   # config 5: uneven 49/49/49/46 pencils
    bench((768, 768, 384), (8, 1), True)
    bench((256, 256, 256), (8, 1), True)
if rank == 0:
    print("WORST", worst, "PASS" if worst <= 1e-12 else "FAIL", flush=True)
dist.destroy_process_group()
