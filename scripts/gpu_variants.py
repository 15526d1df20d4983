"""Time the six stage kernels of the 512^3 (or --n) double r2c+c2r round trip for several builds of the library.
    python scripts/gpu_variants.py lib1.so lib2.so ...      (each in its own subprocess)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch
    import accfft_b200 as acc
    nn = [int(v) for v in os.environ.get("VAR_N", "512").split(",")]
    n3 = tuple(nn) if len(nn) == 3 else (nn[0],) * 3
    n = n3[0]
    kind = os.environ.get("VAR_KIND", "r2c")
    single = os.environ.get("VAR_SINGLE", "0") == "1"
    torch.cuda.set_device(0)
    comm = acc.Comm(0, 1, (1, 1))
    p = acc.Plan(n3, comm, r2c=(kind == "r2c"), single=single)
    rdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    ntot = n3[0] * n3[1] * n3[2]
    x = torch.rand(ntot * (1 if kind == "r2c" else 2), dtype=rdt, device="cuda")
    out = torch.zeros(p.alloc_max // es, dtype=rdt, device="cuda")
    back = torch.zeros_like(x)
    st = torch.cuda.Stream()
    p.set_stream(st.cuda_stream); p.set_blocking(False); p.set_profiling(True)
    def step():
        if kind == "r2c":
            p.execute_r2c(x, out); p.execute_c2r(out, back)
        else:
            p.execute_c2c(acc.ACCFFT_FORWARD, x, out); p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)
    with torch.cuda.stream(st):
        for _ in range(5): step()
    torch.cuda.synchronize(); p.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 30
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(iters): step()
        e1.record(st)
    torch.cuda.synchronize()
    prof = p.profile_read()
    err = float(torch.linalg.vector_norm(back / ntot - x) / torch.linalg.vector_norm(x))
    print(json.dumps({"lib": os.path.basename(os.environ.get("ACCFFT_B200_LIB", "")), "tma": os.environ.get("ACCFFT_B200_TMA", "1"), "n": list(n3), "kind": kind, "kernels": p.stage_kernels(), "single": single,
                      "ms_step": e0.elapsed_time(e1) / iters, "err": err,
                      "stages": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]}}))
    sys.exit(0)
CASES = (("r2c", 512, "0"), ("c2c", 512, "0"), ("r2c", 256, "0"), ("r2c", 512, "1"))
if os.environ.get("VAR_QUICK"):
    CASES = CASES[:1]
if os.environ.get("VAR_CASES"):   # e.g. "r2c:512:1;r2c:768,768,384:0"
    CASES = tuple(tuple(c.split(":")) for c in os.environ["VAR_CASES"].split(";"))
for spec in sys.argv[1:]:
    # "path/to/lib.so" or "path/to/lib.so:ENV=VAL,ENV2=VAL2"
    lib, _, extra = spec.partition(":")
    for kind, n, single in CASES:
        env = dict(os.environ, ACCFFT_B200_LIB=os.path.abspath(lib), VAR_N=str(n), VAR_KIND=kind, VAR_SINGLE=single)
        for kv in filter(None, extra.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or ("FAILED " + lib + " " + r.stderr[-400:]), flush=True)
