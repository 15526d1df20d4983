"""Per-stage device times of a single-GPU plan (event pairs around every launch, accfftb_plan_set_profiling):
    python scripts/stage_profile.py [n0 n1 n2 double|single r2c|c2c] ...      default: the BASELINE sizes
Prints ms per stage, algorithmic GB/s (each pass reads + writes its array once) and the fraction of the measured HBM peak;
then the operator timings (laplace, grad per-axis, grad 3-D) for r2c plans."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import accfft_b200 as acc

try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6550.0


def run(n, single, r2c, iters=10):
    dev = torch.device("cuda", 0)
    comm = acc.Comm(0, 1, (1, 1))
    p = acc.Plan(n, comm, r2c=r2c, single=single)
    rdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    cnt = n[0] * n[1] * n[2] * (1 if r2c else 2)
    x = torch.rand(cnt, dtype=rdt, device=dev) * 2 - 1
    out = torch.zeros(p.alloc_max // es, dtype=rdt, device=dev)
    back = torch.zeros_like(x)
    st = torch.cuda.Stream()
    p.set_stream(st.cuda_stream)
    p.set_blocking(False)

    def step():
        if r2c:
            p.execute_r2c(x, out)
            p.execute_c2r(out, back)
        else:
            p.execute_c2c(acc.ACCFFT_FORWARD, x, out)
            p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)

    with torch.cuda.stream(st):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):   # the round trip is timed WITHOUT per-launch events (they cost a few us per launch)
        e0.record(st)
        for _ in range(iters):
            step()
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    p.set_profiling(True)
    p.profile_read()
    with torch.cuda.stream(st):
        for _ in range(iters):
            step()
    torch.cuda.synchronize()
    prof = p.profile_read()
    p.set_profiling(False)
    n2c = n[2] // 2 + 1 if r2c else n[2]
    cs = 2 * es
    R = cnt * es
    C = n[0] * n[1] * n2c * cs
    algo = {"fwd_z": R + C if r2c else 2 * C, "fwd_y": 2 * C, "fwd_x": 2 * C, "inv_x": 2 * C, "inv_y": 2 * C, "inv_z": C + R if r2c else 2 * C}
    tot = sum(algo.values())
    # fused local pair (csrc/fused.cuh): one kernel, profiled in the slot of its first pass
    order = ["fwd_z", "fwd_y", "fwd_x", "inv_x", "inv_y", "inv_z"]
    if p.fused_info(False) and prof["fwd_y"][1] == 0:
        algo["fwd_zy"] = algo["fwd_z"] + algo["fwd_y"]
        prof["fwd_zy"] = prof["fwd_z"]
        order[0:2] = ["fwd_zy"]
    if p.fused_info(True) and prof["inv_z"][1] == 0:
        algo["inv_yz"] = algo["inv_y"] + algo["inv_z"]
        prof["inv_yz"] = prof["inv_y"]
        order[-2:] = ["inv_yz"]
    print("%s %s %s: round trip %.3f ms, %.0f GB/s algorithmic = %.3f of %.0f; kernels %s" %
          ("x".join(map(str, n)), "single" if single else "double", "r2c" if r2c else "c2c", ms, tot / ms / 1e6, tot / ms / 1e6 / PEAK, PEAK,
           p.stage_kernels()))
    print("   fused pair: forward %s, inverse %s" % (p.fused_info(False), p.fused_info(True)))
    for k in order:
        t, c = prof[k]
        if c:
            tp = t / iters   # per step (a stage may run as several chunk launches)
            print("   %-6s %.4f ms (%d launches)  %6.0f GB/s  %.3f" % (k, tp, c // iters, algo[k] / tp / 1e6, algo[k] / tp / 1e6 / PEAK))
    if r2c:
        la = torch.zeros_like(x)
        g = [torch.zeros_like(x) for _ in range(3)]
        for name, fn, passes in (("laplace (fused symbol)", lambda: p.laplace(la, x), tot),
                                 ("grad per-axis", lambda: p.grad(g[0], g[1], g[2], x), 3 * 2 * R),
                                 ("grad 3-D scheme", lambda: p.grad_slow(g[0], g[1], g[2], x), 2 * tot),
                                 ("divergence per-axis", lambda: p.divergence(la, g[0], g[1], g[2]), 3 * 2 * R + 2 * 2 * R)):
            with torch.cuda.stream(st):
                fn()
                torch.cuda.synchronize()
                e0.record(st)
                for _ in range(5):
                    fn()
                e1.record(st)
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 5
            print("   %-24s %.3f ms  (%.0f GB/s of its algorithmic traffic, %.3f of peak)" % (name, t, passes / t / 1e6, passes / t / 1e6 / PEAK))
    p.destroy()
    del x, out, back
    torch.cuda.empty_cache()


if __name__ == "__main__":
    a = sys.argv[1:]
    if a:
        cases = []
        while a:
            cases.append(((int(a[0]), int(a[1]), int(a[2])), a[3] == "single", a[4] == "r2c"))
            a = a[5:]
    else:
        cases = [((512, 512, 512), False, True), ((256, 256, 256), False, True), ((512, 512, 512), True, True),
                 ((1024, 1024, 1024), True, True), ((512, 512, 512), False, False), ((768, 768, 384), False, True)]
    for n, single, r2c in cases:
        run(n, single, r2c)
