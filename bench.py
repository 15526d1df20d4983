#!/usr/bin/env python
"""bench.py -- headline benchmark of accfft_b200 (BASELINE.json): 3D r2c+c2r round trip at 512^3 double,
GFLOP/s with FLOPs = 5 N log2 N per transform (N = N0 N1 N2; the same count for r2c and c2r), on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one forward r2c + one inverse c2r of the whole grid (strong scaling: the 512^3 grid is split
over the N ranks as a P0 x P1 pencil grid).  Prints ONE JSON line on rank 0.

  value     device-timed (CUDA events on the plan's stream, max over ranks), inputs resident in HBM
  e2e       same metric through the public API with HOST buffers: pinned H2D of the input field and D2H of
            the round-trip field inside the timed region, every step
  roofline  the dominant stage kernel: algorithmic HBM bytes per launch / its mean launch duration (event
            pairs recorded around every launch inside the timed region) vs the measured HBM copy peak
  cpu_baseline / --impl reference: the CPU oracle port (pocketfft via scipy.fft, all host threads) -- the
            reference's own FFTW+MPI path cannot be built in this image (no MPI, no FFTW)
"""
import argparse
import json
import math
import os
import sys
import threading
import time

# plan stream + copy lanes + the e2e copy streams + NCCL's own: keep every stream on its own hardware queue (default: 8)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Process grids (the caller's c_dims in the reference: any P0 x P1 = N works).  Defaults = the fastest measured here for
# the headline 512^3 grid (profiles/mgpu8_r14.log: 8x1 0.696 ms, 4x2 0.764, 2x4 0.771, 1x8 0.765 per round trip;
# profiles/mgpu4_r15.log: 4x1 1.111, 2x2 1.134, 1x4 1.153): a slab
# split has ONE exchange per transform and three local stages; --grid P0xP1 selects any other (e.g. the 2x4 pencil
# grid BASELINE.json names for its 1024^3 / uneven configs).
GRIDS = {1: (1, 1), 2: (2, 1), 4: (4, 1), 8: (8, 1)}
NVLINK_GBS = 770.0   # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md:24-29; 900 nominal)


def flops_per_transform(n):
    N = n[0] * n[1] * n[2]
    return 5.0 * N * math.log2(N)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.stop_flag = index, period, False
        self.sm, self.reasons, self.sm_max, self.err = [], set(), None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
                nv.nvmlClocksThrottleReasonApplicationsClocksSetting: "applications_clocks_setting",
            }
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "error": self.err}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_roundtrip(n, workers, x=None):
    from oracle import accfft_oracle as orc
    if x is None:
        x = orc.cpu_input(n)
    t, _w = orc.cpu_roundtrip_seconds(n, workers=workers, x=x)
    return t


THREAD_ENV = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "DUCC0_NUM_THREADS",
              "VECLIB_MAXIMUM_THREADS", "OMP_THREAD_LIMIT")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which pins scipy's FFT thread pool to ONE thread
    # (round 1: 13.9 GFLOP/s under torchrun vs 179 stand-alone on the same box).  The reference arm is "all the host
    # threads it can use": drop every thread cap BEFORE numpy / scipy are imported.
    capped = {k: os.environ.pop(k) for k in THREAD_ENV if k in os.environ}
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))   # and undo any inherited CPU pinning
    except Exception:
        pass
    from oracle import accfft_oracle as orc
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n_full = (args.n, args.n, args.n)
    n_small = (min(256, args.n),) * 3
    cpu_roundtrip((64, 64, 64), cores)                      # spin up the pocketfft thread pool
    t_small = cpu_roundtrip(n_small, cores)
    budget = 150.0
    total = args.steps + args.warmup
    scale = (args.n / n_small[0]) ** 3 * 1.15
    if t_small * scale * total <= budget:
        n = n_full
    elif t_small * total <= budget:
        n = n_small
    else:
        n = (128, 128, 128)
    x = orc.cpu_input(n)
    for _ in range(args.warmup):
        cpu_roundtrip(n, cores, x)
    c0, w0 = time.process_time(), time.perf_counter()
    dt = sum(cpu_roundtrip(n, cores, x) for _ in range(args.steps)) / max(args.steps, 1)
    busy = (time.process_time() - c0) / max(time.perf_counter() - w0, 1e-9)   # CPU seconds per wall second = threads kept busy
    gf = 2 * flops_per_transform(n) / dt / 1e9
    sample = "%d^3 double r2c+c2r round trip per step (scipy.fft pocketfft, workers=%d, %.1f threads busy on average)" % (
        n[0], cores, busy)
    if cores > 2 and busy < 1.5 and n[0] >= 256:   # (tiny grids do not keep a thread pool busy)
        raise SystemExit("reference arm ran on %.1f threads of %d: a thread cap is still in force (%r)" % (busy, cores, capped))
    line = {
        "impl": "reference", "metric": METRIC, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%d^3 double r2c+c2r round trip (forward + inverse = 1 step)" % n[0], "sample": sample,
                   "requested": "%d^3" % args.n, "thread_caps_removed": capped, "threads_busy": busy,
                   "note": "CPU oracle port (scipy pocketfft): the stand-in for FFTW+MPI, which cannot be built here (no MPI, "
                           "no FFTW); oracle/_ref holds the reference's own decomposition / transpose / operator code built "
                           "against a definitional-DFT fftw3.h shim -- a parity pin, not a timing baseline"},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0



# ---------------------------------------------------------------------------------------------------
# parity inside the bench (N > 1): every rank checks ITS blocks against the oracle (tests/ holds the full suite; this
# puts multi-GPU forward-spectrum / operator parity on real GPUs into the driver-visible record)
# ---------------------------------------------------------------------------------------------------
PENCIL = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}     # the pencil grids BASELINE.json names (2x4 at 8 GPUs)
PARITY_SIZES = [(128, 96, 160), (5, 6, 6)]               # uneven pencils; (5, 6, 6) leaves ranks short or empty


def parity_case(acc, orc, torch, dist, dev, n, dims, world, rank):
    """-> {check: max over ranks of the relative L2 error vs the oracle} for one grid; all ranks must call"""
    import numpy as np
    comm = acc.accfft_create_comm(list(dims), rank=rank, nprocs=world)
    p = acc.Plan(n, comm, r2c=True, single=False, flags=acc.ACCFFT_ESTIMATE)
    x = orc.synthetic_real(n, 5)
    X = orc.forward_r2c(x)
    loc = orc.local_input(x, n, dims, comm.coords)
    xd = torch.from_numpy(np.ascontiguousarray(loc)).to(dev) if loc.size else torch.zeros(2, dtype=torch.float64, device=dev)
    out = torch.zeros(max(p.alloc_max // 8, 2), dtype=torch.float64, device=dev)
    back = torch.zeros_like(xd)
    errs = {"forward": 0.0, "round_trip": 0.0, "laplace": 0.0, "grad": 0.0}
    for _ in range(2):
        p.execute_r2c(xd, out)
        p.execute_c2r(out, back)
    osz = p.osize
    cnt = osz[0] * osz[1] * osz[2]
    if cnt:
        got = out.cpu().numpy().view(np.complex128)[:cnt].reshape(osz)
        errs["forward"] = orc.rel_l2(got, orc.local_output(X, n, dims, comm.coords, r2c=True))
    la = torch.zeros_like(xd)
    g = [torch.zeros_like(xd) for _ in range(3)]
    p.laplace(la, xd)
    p.grad(g[0], g[1], g[2], xd)
    if loc.size:
        errs["round_trip"] = orc.rel_l2(back.cpu().numpy().reshape(loc.shape) / x.size, loc)
        errs["laplace"] = orc.rel_l2(la.cpu().numpy().reshape(loc.shape), orc.local_input(orc.laplace(x), n, dims, comm.coords))
        errs["grad"] = max(orc.rel_l2(g[ax].cpu().numpy().reshape(loc.shape), orc.local_input(orc.grad(x, ax), n, dims, comm.coords))
                           for ax in range(3))
    mode = p.exchange_mode()["mode"]
    p.destroy()
    keys = sorted(errs)
    t = torch.tensor([errs[k] for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {k: float(v) for k, v in zip(keys, t.tolist())}, mode


def run_parity(acc, torch, dist, dev, world, rank, run_dims):
    from oracle import accfft_oracle as orc
    grids = [tuple(run_dims)]
    if tuple(PENCIL.get(world, run_dims)) not in grids:
        grids.append(tuple(PENCIL[world]))
    out, worst = {}, 0.0
    for dims in grids:
        for n in PARITY_SIZES:
            errs, mode = parity_case(acc, orc, torch, dist, dev, n, dims, world, rank)
            out["%dx%dx%d on %dx%d (%s)" % (n[0], n[1], n[2], dims[0], dims[1], mode)] = errs
            worst = max(worst, max(errs.values()))
    return out, worst


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs 3-5 (the non-headline configurations), time-boxed: a few iterations each, on the pencil grid
# BASELINE names and on the slab grid; each with its own round-trip guard and roofline fraction
# ---------------------------------------------------------------------------------------------------
def time_plan(torch, dist, dev, world, fn, iters, warm=2):
    st = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def config_entry(acc, torch, dist, dev, world, rank, name, n, dims, r2c, single, peak, operators=False, iters=5):
    """one BASELINE config on one grid: round trip ms / TFLOP/s / fraction of max(HBM, NVLink) roofline (+ operators)"""
    comm = acc.accfft_create_comm(list(dims), rank=rank, nprocs=world)
    p = acc.Plan(n, comm, r2c=r2c, single=single)
    rdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    isz, osz = p.isize, p.osize
    cnt = isz[0] * isz[1] * isz[2] * (1 if r2c else 2)
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    x = torch.rand(max(cnt, 2), dtype=rdt, device=dev, generator=gen) * 2 - 1
    out = torch.zeros(max(p.alloc_max // es, 2), dtype=rdt, device=dev)
    back = torch.zeros_like(x)
    p.set_stream(torch.cuda.current_stream().cuda_stream)
    p.set_blocking(False)

    def step():
        if r2c:
            p.execute_r2c(x, out)
            p.execute_c2r(out, back)
        else:
            p.execute_c2c(acc.ACCFFT_FORWARD, x, out)
            p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)

    ms = time_plan(torch, dist, dev, world, step, iters)
    N = n[0] * n[1] * n[2]
    err = float(torch.linalg.vector_norm(back[:cnt] / N - x[:cnt]) / torch.linalg.vector_norm(x[:cnt])) if cnt else 0.0
    t = torch.tensor([err], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err = float(t.item())
    # roofline of this rank's share: every axis pass reads and writes its array once; the exchanged part crosses NVLink
    n2c = n[2] // 2 + 1 if r2c else n[2]
    cs = 2 * es
    R = isz[0] * isz[1] * isz[2] * (es if r2c else cs)
    C0 = isz[0] * isz[1] * n2c * cs
    C1 = isz[0] * n[1] * osz[2] * cs
    C2 = n[0] * osz[1] * osz[2] * cs
    hbm = 2 * (R + C0 + 2 * C1 + 2 * C2)
    P0, P1 = dims
    nvl = 2 * (C0 * (P1 - 1) / P1 + C1 * (P0 - 1) / P0)
    t_roof = max(hbm / (peak * 1e9), nvl / (NVLINK_GBS * 1e9)) * 1e3
    e = {"grid": "%dx%d" % dims, "exchange": p.exchange_mode()["mode"], "ms": ms, "tflops": 2 * flops_per_transform(n) / (ms * 1e-3) / 1e12,
         "t_roof_ms": t_roof, "frac_of_roofline": t_roof / ms, "round_trip_rel_l2": err,
         "tolerance": 1e-5 if single else 1e-12, "ok": err <= (1e-5 if single else 1e-12)}
    if operators and r2c:
        la = torch.zeros_like(x)
        g = [torch.zeros_like(x) for _ in range(3)]
        e["laplace_ms"] = time_plan(torch, dist, dev, world, lambda: p.laplace(la, x), iters)
        e["grad_ms"] = time_plan(torch, dist, dev, world, lambda: p.grad(g[0], g[1], g[2], x), iters)
        # laplace = one forward + one inverse with the symbol fused into the X stage: same roofline as the round trip;
        # grad (per-axis scheme): three single-axis derivative passes, each reads and writes the real field once (+ exchanges)
        e["laplace_frac_of_roofline"] = t_roof / e["laplace_ms"]
    p.destroy()
    del x, out, back
    torch.cuda.empty_cache()
    return e


def batched_entry(acc, torch, dist, dev, world, rank, n, dims, nf=3, iters=5):
    """nf fields (a vector field) through ONE set of exchanges (accfftb_execute_*_many) vs nf single-field round trips
    on the same plan; plus the full-3-D gradient / divergence, which batch their components on such a plan, vs a one-field plan"""
    comm = acc.accfft_create_comm(list(dims), rank=rank, nprocs=world)
    p = acc.Plan(n, comm, r2c=True, single=False, howmany=nf)
    isz = p.isize
    cnt = max(isz[0] * isz[1] * isz[2], 2)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    xs = [torch.rand(cnt, dtype=torch.float64, device=dev, generator=gen) * 2 - 1 for _ in range(nf)]
    outs = [torch.zeros(max(p.alloc_max // 8, 2), dtype=torch.float64, device=dev) for _ in range(nf)]
    backs = [torch.zeros_like(x) for x in xs]
    p.set_stream(torch.cuda.current_stream().cuda_stream)
    p.set_blocking(False)

    def seq():
        for f in range(nf):
            p.execute_r2c(xs[f], outs[f])
            p.execute_c2r(outs[f], backs[f])

    def many():
        p.execute_r2c_many(xs, outs)
        p.execute_c2r_many(outs, backs)

    ms_seq = time_plan(torch, dist, dev, world, seq, iters)
    ref = [b.clone() for b in backs]
    for b in backs:
        b.zero_()
    ms_many = time_plan(torch, dist, dev, world, many, iters)
    same = all(bool(torch.equal(a, b)) for a, b in zip(ref, backs))
    div = torch.zeros_like(xs[0])
    g = [torch.zeros_like(xs[0]) for _ in range(3)]
    e = {"grid": "%dx%d" % dims, "fields": nf, "ms_sequential": ms_seq, "ms_batched": ms_many, "speedup": ms_seq / ms_many,
         "bitwise_equal": same, "ok": same}
    if nf >= 3:
        e["grad3d_batched_ms"] = time_plan(torch, dist, dev, world, lambda: p.grad_slow(g[0], g[1], g[2], xs[0]), iters)
        e["div3d_batched_ms"] = time_plan(torch, dist, dev, world, lambda: p.divergence_slow(div, xs[0], xs[1], xs[2]), iters)
    p.destroy()
    if nf >= 3:
        p1 = acc.Plan(n, comm, r2c=True, single=False)
        p1.set_stream(torch.cuda.current_stream().cuda_stream)
        p1.set_blocking(False)
        e["grad3d_ms"] = time_plan(torch, dist, dev, world, lambda: p1.grad_slow(g[0], g[1], g[2], xs[0]), iters)
        e["div3d_ms"] = time_plan(torch, dist, dev, world, lambda: p1.divergence_slow(div, xs[0], xs[1], xs[2]), iters)
        p1.destroy()
    del xs, outs, backs, ref
    torch.cuda.empty_cache()
    return e


def run_configs(acc, torch, dist, dev, world, rank, peak):
    slab, pencil = (world, 1), tuple(PENCIL.get(world, (world, 1)))
    grids = [pencil] + ([slab] if slab != pencil else [])
    specs = [("config3: 512^3 double c2c forward + backward", (512, 512, 512), False, False, False),
             ("config4: 1024^3 single r2c + c2r, fused laplace / grad operators", (1024, 1024, 1024), True, True, True),
             ("config5: 768x768x384 double r2c + c2r (mixed radix 2/3, uneven pencils)", (768, 768, 384), True, False, False)]
    out = {}
    for name, n, r2c, single, ops in specs:
        out[name] = []
        for dims in grids:
            try:
                out[name].append(config_entry(acc, torch, dist, dev, world, rank, name, n, dims, r2c, single, peak, operators=ops))
            except Exception as ex:   # a config that does not fit (memory) is reported, not fatal for the headline line
                out[name].append({"grid": "%dx%d" % dims, "error": repr(ex)[:200]})
                torch.cuda.empty_cache()
    name = "batched fields: 3 x 512^3 double r2c + c2r through one set of exchanges (accfftb_execute_*_many) vs one by one"
    out[name] = []
    for dims in grids:
        try:
            out[name].append(batched_entry(acc, torch, dist, dev, world, rank, (512, 512, 512), dims))
        except Exception as ex:
            out[name].append({"grid": "%dx%d" % dims, "error": repr(ex)[:200]})
            torch.cuda.empty_cache()
    return out


METRIC = "3D r2c+c2r GFLOP/s (5N*log2N per transform) at 512^3 double"


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=512, help="grid edge (default: the headline 512)")
    ap.add_argument("--grid", default=None, help="process grid P0xP1 (default: see GRIDS)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity block (N > 1)")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 3-5 block (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import accfft_b200 as acc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    n = (args.n, args.n, args.n)
    dims = list(GRIDS.get(world, (0, 0)))
    if args.grid:
        dims = [int(v) for v in args.grid.lower().split("x")]
        if dims[0] * dims[1] != world:
            raise SystemExit("--grid %s does not match %d ranks" % (args.grid, world))
    peak, peak_src = measured_peaks()
    # ---- multi-GPU parity, BEFORE anything is timed: no value is printed if a block of any rank is off ----------------
    parity = None
    if world > 1 and not args.no_parity:
        parity, worst = run_parity(acc, torch, dist, dev, world, rank, dims)
        if not worst <= 1e-12:
            raise SystemExit("multi-GPU parity failed (worst relative L2 %.3e > 1e-12): %s" % (worst, json.dumps(parity)))
        parity = {"tolerance": 1e-12, "worst": worst, "checks": "forward spectrum / round trip / laplace / grad, per rank vs the "
                  "oracle, max over ranks", "cases": parity}
    comm = acc.accfft_create_comm(dims, rank=rank, nprocs=world)
    plan = acc.accfft_plan_dft_3d_r2c_gpu(n, comm)
    isz, osz = plan.isize, plan.osize
    n_in = isz[0] * isz[1] * isz[2]
    gen = torch.Generator(device=dev).manual_seed(20240607 + rank)
    x = torch.rand(max(n_in, 1), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    out = torch.zeros(max(plan.alloc_max // 8, 2), dtype=torch.float64, device=dev)
    back = torch.zeros_like(x)
    stream = torch.cuda.Stream()
    plan.set_stream(stream.cuda_stream)
    plan.set_blocking(False)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        plan.execute_r2c(x, out)
        plan.execute_c2r(out, back)

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
    sync_all()
    # correctness guard inside the bench: the round trip must reproduce the input (value is void otherwise)
    N = n[0] * n[1] * n[2]
    rt_err = float(torch.linalg.vector_norm(back / N - x) / torch.linalg.vector_norm(x)) if n_in else 0.0
    if not rt_err <= 1e-12:
        raise SystemExit("round-trip error %g exceeds 1e-12: refusing to report a number" % rt_err)

    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = plan.launch_count
    copies0 = plan.copy_count
    sync_all()
    # the timed region: EXACTLY args.steps steps between two events on the plan's stream, nothing else enqueued
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = plan.launch_count - launches0
    copies = plan.copy_count - copies0
    # per-kernel durations for the roofline: the same args.steps steps once more, this time with an event pair around every
    # launch (accfftb_plan_set_profiling).  Kept out of the region above because the 12 extra event records per step keep
    # consecutive kernels from running back to back (measured: +3 % on the step).
    # The pass starts like the timed region did -- after a pause and the same warm-up -- because the board's power limiter
    # (clocks `reasons`: sw_power_cap) takes ~0.1 s of continuous load to pull the SM clock down: a pass that followed the timed
    # region immediately would run every kernel under a lower clock than the region itself saw (measured: 8 % slower).
    time.sleep(1.0)
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
    sync_all()
    plan.set_profiling(True)
    plan.profile_read()
    # ... in up to four chunks of >= 5 steps; a stage's duration is the MEDIAN of its chunk averages, so that one stalled launch
    # (seen once: a single 3 ms hiccup in 20 launches of one stage) cannot pose as that kernel's speed
    nchunk = max(1, min(4, args.steps // 5))
    sizes = [args.steps // nchunk + (1 if c < args.steps % nchunk else 0) for c in range(nchunk)]
    chunk_avgs, counts = {}, {}
    for nsteps in sizes:
        with torch.cuda.stream(stream):
            for _ in range(nsteps):
                step()
        sync_all()
        for k, (tms, cnt) in plan.profile_read().items():
            if cnt:
                chunk_avgs.setdefault(k, []).append(tms / cnt)
                counts[k] = counts.get(k, 0) + cnt
    prof = {k: (0.0, 0) for k in plan.PROFILE_SLOTS}
    for k, avgs in chunk_avgs.items():
        srt = sorted(avgs)
        med = srt[len(srt) // 2] if len(srt) % 2 else 0.5 * (srt[len(srt) // 2 - 1] + srt[len(srt) // 2])
        prof[k] = (med * counts[k], counts[k])
    plan.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 2 * flops_per_transform(n) / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: HOST buffers in, HOST buffers out -------------------------------------------------------------
    # Every step copies its input field from pinned host memory to the device, runs r2c + c2r through the public
    # API and copies the round-trip field back to pinned host memory.  Steps are software-pipelined over two device
    # buffer sets and three streams (H2D of step k+1 and D2H of step k-1 overlap the transforms of step k: PCIe is
    # full duplex and the copy engines are independent), the way a streaming caller would drive the library; the
    # timed region covers ALL copies and transforms of the e_steps steps and ends when the last result is on the host.
    e2e = None
    if not args.no_e2e:
        e_steps = max(4, min(args.steps, 12))
        xh = torch.empty(x.shape, dtype=torch.float64).pin_memory()
        xh.copy_(x)
        yh = [torch.empty(back.shape, dtype=torch.float64).pin_memory() for _ in range(2)]
        xd = [x, torch.empty_like(x)]
        bd = [back, torch.empty_like(back)]
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]      # input k landed
        ev_fft = [torch.cuda.Event() for _ in range(2)]     # transforms of step k done (input buffer free, result ready)
        ev_out = [torch.cuda.Event() for _ in range(2)]     # result k on the host (result buffer free)

        def e2e_run(nsteps):
            for k in range(nsteps):
                j = k & 1
                with torch.cuda.stream(s_in):
                    if k >= 2:
                        s_in.wait_event(ev_fft[j])
                    xd[j].copy_(xh, non_blocking=True)
                    ev_in[j].record(s_in)
                with torch.cuda.stream(stream):
                    stream.wait_event(ev_in[j])
                    if k >= 2:
                        stream.wait_event(ev_out[j])
                    plan.execute_r2c(xd[j], out)
                    plan.execute_c2r(out, bd[j])
                    ev_fft[j].record(stream)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_fft[j])
                    yh[j].copy_(bd[j], non_blocking=True)
                    ev_out[j].record(s_out)

        e2e_run(2)
        sync_all()
        ev0.record(stream)
        s_in.wait_event(ev0)
        e2e_run(e_steps)
        stream.wait_stream(s_out)
        ev1.record(stream)
        sync_all()
        e_ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e_ms /= e_steps
        ok = all(bool(torch.allclose(h[:4096] / N, xh[:4096], atol=1e-12)) and
                 bool(torch.allclose(h[-4096:] / N, xh[-4096:], atol=1e-12)) for h in yh) if n_in else True
        if not ok:
            raise SystemExit("e2e round trip does not reproduce the host input: refusing to report a number")
        e2e = {"value": 2 * flops_per_transform(n) / (e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": int(xh.numel() * 8), "d2h_bytes_per_step": int(yh[0].numel() * 8),
               "ms_per_step": e_ms, "steps": e_steps, "checked": ok,
               "how": "pinned host -> device -> r2c -> c2r -> pinned host every step; steps pipelined over 2 buffer sets / 3 streams"}
    clocks = sampler.result()

    # ---- roofline of the dominant stage kernel -----------------------------------------------------------
    n2c = n[2] // 2 + 1
    k2 = osz[2]
    R = n_in * 8
    C0 = isz[0] * isz[1] * n2c * 16            # after Z:  [i0][i1][N2c]
    C1 = isz[0] * n[1] * k2 * 16               # Y pass:   [i0][N1][k2]
    C2 = n[0] * osz[1] * k2 * 16               # X pass:   [N0][k1][k2]
    algo = {"fwd_z": R + C0, "fwd_y": C1 + C1, "fwd_x": C2 + C2, "inv_x": C2 + C2, "inv_y": C1 + C1, "inv_z": C0 + R}
    if world > 1:      # the stage that feeds an exchange writes the NEXT layout's bytes (same totals on even grids)
        algo["fwd_z"] = R + C0
    # bytes that leave this GPU over NVLink with each exchanging stage (peer stores fused into the stage, or copy-engine
    # transfers of what the stage packed)
    nvl = {"fwd_z": C0 * (dims[1] - 1) / dims[1], "fwd_y": C1 * (dims[0] - 1) / dims[0],
           "inv_x": C2 * (dims[0] - 1) / dims[0], "inv_y": C0 * (dims[1] - 1) / dims[1]}
    # The fused local pair (csrc/fused.cuh): Z and Y of a direction run as ONE persistent kernel, profiled in the slot of its
    # first pass; the array between the two passes is a ring of a few planes that stays in L2.  Its ALGORITHMIC bytes stay
    # those of the two passes (every pass reads and writes its array once); the HBM floor of the kernel is lower by
    # the write + read of the intermediate, so "gbs" of a fused stage may exceed the HBM peak -- hbm_min_bytes says by how much.
    fused = {"forward": plan.fused_info(False), "inverse": plan.fused_info(True)}
    hbm_min = {}
    if fused["forward"] and prof.get("fwd_y", (0, 0))[1] == 0:
        algo["fwd_zy"] = algo.pop("fwd_z") + algo.pop("fwd_y")
        hbm_min["fwd_zy"] = R + C1
        prof["fwd_zy"] = prof.pop("fwd_z")
        if "fwd_y" in nvl:
            nvl["fwd_zy"] = nvl.pop("fwd_y")
    if fused["inverse"] and prof.get("inv_z", (0, 0))[1] == 0:
        algo["inv_yz"] = algo.pop("inv_y") + algo.pop("inv_z")
        hbm_min["inv_yz"] = C1 + R
        prof["inv_yz"] = prof.pop("inv_y")
        nvl.pop("inv_y", None)
    xmode = plan.exchange_mode()
    exchange = dict(xmode)
    if xmode["mode"] == "copy_engine":
        # the exchanging stages pack into a local staging buffer (HBM-bound kernels); copy engines move it over NVLink
        # while the next chunk computes: the link traffic belongs to the transfers, not to a kernel
        nvl_step = sum(nvl.values())
        cms, ccnt = prof.get("ce_copy", (0.0, 0))
        exchange.update({"nvlink_out_bytes_per_step": int(nvl_step), "transfers_per_step": ccnt / args.steps,
                         "transfer_busy_ms_per_step": cms / args.steps,
                         "nvlink_out_gbs": (nvl_step / (cms / args.steps * 1e-3) / 1e9) if cms else None,
                         "nvlink_peak_gbs": NVLINK_GBS,
                         "nvlink_frac": (nvl_step / (cms / args.steps * 1e-3) / 1e9 / NVLINK_GBS) if cms else None,
                         "overlap": "transfer of chunk c runs on copy lanes while chunk c+1 computes; busy time = union of "
                                    "the transfer intervals (CUDA events on the copy streams)"})
        nvl = {}
    stages = {}
    for k, (tms, cnt) in prof.items():
        if cnt:
            stages[k] = {"ms": tms / cnt, "launches_per_step": cnt / args.steps}
            if k in algo:
                stages[k]["algo_bytes"] = algo[k]
                stages[k]["gbs"] = algo[k] / (tms / args.steps * 1e-3) / 1e9     # per-step bytes / per-step time
            if k in hbm_min:
                stages[k]["hbm_min_bytes"] = hbm_min[k]
                stages[k]["hbm_min_gbs"] = hbm_min[k] / (tms / args.steps * 1e-3) / 1e9
            if nvl.get(k):
                stages[k]["nvlink_out_bytes"] = int(nvl[k])
                stages[k]["nvlink_out_gbs"] = nvl[k] / (tms / args.steps * 1e-3) / 1e9
    dom = max((k for k in stages if k in algo), key=lambda k: stages[k]["ms"] * stages[k]["launches_per_step"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        key = "%s_n%d_p%d" % (dom, args.n, world)
        traffic = tr.get(key)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["gbs"] / peak, "traffic": traffic,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size, one --set full "
                                  "capture (profiles/traffic.json); not re-measured in this run" if traffic else None,
                "algo_bytes_per_launch": algo[dom] / stages[dom]["launches_per_step"],
                "ms_per_launch": stages[dom]["ms"], "peak_source": peak_src,
                "timing": "CUDA event pair around every launch of a second pass over the same %d steps, on the plan's stream "
                          "(the timed region behind `value` carries no per-launch events); per stage the median of the "
                          "averages of up to four chunks of that pass" % args.steps}
    if dom in hbm_min:
        roofline.update({"hbm_min_bytes_per_launch": hbm_min[dom] / stages[dom]["launches_per_step"],
                         "hbm_min_frac": stages[dom]["hbm_min_gbs"] / peak,
                         "note": "fused Z+Y persistent kernel: algorithmic bytes = both passes read and write their array once "
                                 "(SURVEY 8d); the array between the passes lives in an L2-resident ring, so the kernel's HBM "
                                 "floor is hbm_min_bytes and `frac` may exceed 1 -- hbm_min_frac is the fraction of the HBM peak "
                                 "its unavoidable traffic reaches"})
    if "nvlink_out_gbs" in stages[dom]:
        # the dominant kernel stores to peers (exchange fused into the FFT kernel): its ceiling is the link, not HBM
        roofline.update({"bound": "nvlink", "achieved": stages[dom]["nvlink_out_gbs"], "peak": NVLINK_GBS,
                         "frac": stages[dom]["nvlink_out_gbs"] / NVLINK_GBS, "hbm_gbs": stages[dom]["gbs"],
                         "hbm_frac": stages[dom]["gbs"] / peak,
                         "peak_source": "measured peer-copy bandwidth per direction (B200_PROFILING.md)",
                         "note": "achieved = bytes this stage stores into peers' buffers / its duration"})
    # whole-step view: sum of algorithmic bytes of the six passes vs the step time
    total_algo = sum(algo.values())
    # roofline of the whole step: perfect-overlap model  T = max(HBM bytes / HBM peak, NVLink bytes out / link peak)
    # NVLink bytes leaving this GPU per transform = the part of each exchanged array that lands on OTHER ranks
    P0, P1 = dims
    nvl_out = 2 * (C0 * (P1 - 1) / P1 + C1 * (P0 - 1) / P0)
    t_hbm = total_algo / (peak * 1e9) * 1e3
    t_nvl = nvl_out / (NVLINK_GBS * 1e9) * 1e3
    step_roof = {"algo_bytes_per_step": total_algo, "gbs": total_algo / (ms_per_step * 1e-3) / 1e9,
                 "frac_of_hbm_peak": total_algo / (ms_per_step * 1e-3) / 1e9 / peak,
                 "nvlink_bytes_out_per_step": int(nvl_out), "t_hbm_ms": t_hbm, "t_nvlink_ms": t_nvl,
                 "t_roof_ms": max(t_hbm, t_nvl), "frac_of_roofline": max(t_hbm, t_nvl) / ms_per_step,
                 "model": "max(HBM bytes / %.0f GB/s, NVLink bytes out / %.0f GB/s), rank 0's share" % (peak, NVLINK_GBS)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        from oracle import accfft_oracle as orc
        cpu_roundtrip((64, 64, 64), cores)
        xc = orc.cpu_input(n)
        t_cpu = cpu_roundtrip(n, cores, xc)
        reps = 1
        while reps < 3 and t_cpu * (reps + 1) < 20:
            t_cpu = min(t_cpu, cpu_roundtrip(n, cores, xc))
            reps += 1
        del xc
        cpu_baseline = {"value": 2 * flops_per_transform(n) / t_cpu / 1e9, "unit": "GFLOP/s", "cores": cores,
                        "kind": "port", "seconds_per_round_trip": t_cpu,
                        "sample": "best of %d full %d^3 double r2c+c2r round trip(s), scipy.fft pocketfft workers=%d "
                                  "(stand-in for FFTW+MPI, which cannot be built here)" % (reps, args.n, cores)}

    # ---- BASELINE configs 3-5 (N > 1): after the headline plan is gone, so that their buffers fit next to nothing ----
    stage_kernels = plan.stage_kernels()
    plan.destroy()
    del x, out, back
    torch.cuda.empty_cache()
    configs = None
    if world > 1 and not args.no_configs and args.n == 512:
        configs = run_configs(acc, torch, dist, dev, world, rank, peak)
        bad = [(k, c) for k, v in configs.items() for c in v if c.get("ok") is False]
        if bad:
            raise SystemExit("a BASELINE config failed its round-trip guard: %s" % json.dumps(bad))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%d^3 double r2c+c2r round trip (forward + inverse = 1 step)" % args.n,
                       "grid": "%dx%d" % (dims[0], dims[1]),
                       "flops": "2 x 5 N log2 N per step", "l2": "working set %.2f GB per rank >> 126 MB L2, no flush" %
                       ((R + 2 * C0) / 1e9), "round_trip_rel_l2": rt_err},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "stages": stages, "step_roofline": step_roof,
            "stage_kernels": stage_kernels, "fused_pair": fused, "exchange": exchange, "gpu_copies": int(copies),
            "parity": parity, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
