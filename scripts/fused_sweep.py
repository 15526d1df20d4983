"""Sweep of the fused local Z / Y pair (csrc/fused.cuh) on one GPU: round-trip time and per-slot times of a plan for a list of
(ACCFFT_B200_FUSED, _CP, _LAG, _NS, _FLAGS) settings (read at plan creation).
    python scripts/fused_sweep.py [n0 n1 n2 double|single r2c|c2c] [quick]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import accfft_b200 as acc

PEAK = 6550.0


REF = {}


def run(n, single, r2c, env, iters=10):
    for k in ("FUSED", "FUSED_CP", "FUSED_LAG", "FUSED_NS", "FUSED_FLAGS"):
        os.environ.pop("ACCFFT_B200_" + k, None)
    for k, v in env.items():
        os.environ["ACCFFT_B200_" + k] = str(v)
    dev = torch.device("cuda", 0)
    p = acc.Plan(n, acc.Comm(0, 1, (1, 1)), r2c=r2c, single=single)
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
    err = float(torch.linalg.vector_norm(back / (n[0] * n[1] * n[2]) - x) / torch.linalg.vector_norm(x))
    # bitwise against the two-kernel path on the same input (first call of the sweep = FUSED 0)
    torch.manual_seed(1)
    xr = torch.rand(cnt, dtype=rdt, device=dev) * 2 - 1
    o2, b2 = torch.zeros_like(out), torch.zeros_like(back)
    with torch.cuda.stream(st):
        if r2c:
            p.execute_r2c(xr, o2)
            p.execute_c2r(o2, b2)
        else:
            p.execute_c2c(acc.ACCFFT_FORWARD, xr, o2)
            p.execute_c2c(acc.ACCFFT_BACKWARD, o2, b2)
    torch.cuda.synchronize()
    key = (n, single, r2c)
    if key not in REF:
        REF[key] = (o2.clone(), b2.clone())
        same = "ref"
    else:
        same = "fwd %s inv %s" % ("same" if torch.equal(o2, REF[key][0]) else "DIFF(%d)" % int((o2 != REF[key][0]).sum()),
                                  "same" if torch.equal(b2, REF[key][1]) else "DIFF(%d)" % int((b2 != REF[key][1]).sum()))
    del xr, o2, b2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
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
    fi, ii = p.fused_info(False), p.fused_info(True)
    p.destroy()
    C = n[0] * n[1] * (n[2] // 2 + 1 if r2c else n[2]) * 2 * es
    R = n[0] * n[1] * n[2] * es
    algo = 2 * (R + 5 * C) if r2c else 12 * C
    stages = " ".join("%s %.3f" % (k, v[0] / v[1]) for k, v in prof.items() if v[1])
    print("%-44s %.3f ms  %5.0f GB/s algo (%.3f of %d)  rt-err %.1e %s | %s | fwd %s inv %s" % (
        env, ms, algo / ms / 1e6, algo / ms / 1e6 / PEAK, PEAK, err, same, stages,
        None if fi is None else (fi["cp"], fi["lag"], fi["ns"], fi["ring_planes"], fi["grid"]),
        None if ii is None else (ii["cp"], ii["lag"], ii["ns"], ii["ring_planes"], ii["grid"])), flush=True)
    del x, out, back
    torch.cuda.empty_cache()


if __name__ == "__main__":
    a = sys.argv[1:]
    quick = "quick" in a
    a = [v for v in a if v != "quick"]
    n = tuple(int(v) for v in a[:3]) if len(a) >= 3 else (512, 512, 512)
    single = len(a) > 3 and a[3] == "single"
    r2c = not (len(a) > 4 and a[4] == "c2c")
    print("# %s %s %s" % (n, "single" if single else "double", "r2c" if r2c else "c2c"))
    envs = [{"FUSED": 0}, {"FUSED": 1}, {"FUSED": 1, "FUSED_FLAGS": 1}, {"FUSED": 1, "FUSED_FLAGS": 3}, {"FUSED": 1, "FUSED_FLAGS": 4},
            {"FUSED": 1, "FUSED_FLAGS": 12}]
    if not quick:
        envs += [dict(e, FUSED=1) for e in [{"FUSED_CP": 1, "FUSED_LAG": 8, "FUSED_NS": 12}, {"FUSED_CP": 4}, {"FUSED_CP": 8}, {"FUSED_CP": 2, "FUSED_LAG": 8, "FUSED_NS": 16},
                 {"FUSED_CP": 4, "FUSED_LAG": 3, "FUSED_NS": 6}, {"FUSED_CP": 4, "FUSED_LAG": 4, "FUSED_NS": 8}]]
    for env in envs:
        try:
            run(n, single, r2c, env)
        except Exception as e:  # pragma: no cover
            print(env, "FAILED", repr(e), flush=True)
