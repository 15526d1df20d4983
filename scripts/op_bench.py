"""Operator / single-axis timings on one GPU (CUDA events on the plan's stream, warm-up 3, 10 timed calls each).
Prints one JSON line per measurement: the reference's two operator schemes (per-axis = accfft_grad_gpu, full 3-D =
accfft_grad_gpu_slow), laplace, and the six single-axis transforms, with the algorithmic HBM bytes of each.
Usage: python scripts/op_bench.py [n=512] [single=0]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import accfft_b200 as acc


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    single = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    tdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    shape = (n, n, n)
    comm = acc.accfft_create_comm([1, 1], rank=0, nprocs=1)
    plan = acc.Plan(shape, comm, r2c=True, single=single)
    x = torch.rand(shape, dtype=tdt, device=dev) * 2 - 1
    outs = [torch.zeros_like(x) for _ in range(4)]
    spec = torch.zeros(plan.alloc_max // es, dtype=tdt, device=dev)
    stream = torch.cuda.Stream()
    plan.set_stream(stream.cuda_stream)
    plan.set_blocking(False)
    R = n ** 3 * es
    C = n * n * (n // 2 + 1) * 2 * es

    warm = int(os.environ.get("OP_BENCH_WARMUP", "3"))

    def timeit(name, fn, algo_bytes, reps=int(os.environ.get("OP_BENCH_REPS", "10"))):
        with torch.cuda.stream(stream):
            for _ in range(warm):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"op": name, "n": n, "single": single, "ms": round(ms, 4), "algo_GB": round(algo_bytes / 1e9, 3),
                          "GBps": round(algo_bytes / ms / 1e6, 1)}), flush=True)

    timeit("laplace (r2c+c2r, symbol fused)", lambda: plan.laplace(outs[0], x), 2 * (R + 5 * C))
    timeit("grad per-axis (x,y fused 1 pass; z 2 passes)", lambda: plan.grad(outs[0], outs[1], outs[2], x), 2 * R + 2 * R + 2 * (R + C))
    timeit("grad slow (4 full 3-D transforms)", lambda: plan.grad_slow(outs[0], outs[1], outs[2], x), 4 * (R + 5 * C))
    timeit("grad x only", lambda: plan.grad(outs[0], None, None, x, xyz=1), 2 * R)
    timeit("grad y only", lambda: plan.grad(None, outs[1], None, x, xyz=2), 2 * R)
    timeit("grad z only", lambda: plan.grad(None, None, outs[2], x, xyz=4), 2 * (R + C))
    timeit("divergence per-axis", lambda: plan.divergence(outs[3], outs[0], outs[1], outs[2]), 2 * R + 3 * R + 2 * (R + C) + R)
    timeit("divergence slow", lambda: plan.divergence_slow(outs[3], outs[0], outs[1], outs[2]), 4 * (R + 5 * C))
    for ax in (0, 1, 2):
        timeit("r2c axis %s" % "xyz"[ax], lambda: plan.execute_r2c_axis(ax, x, spec), R + C)
        timeit("c2r axis %s" % "xyz"[ax], lambda: plan.execute_c2r_axis(ax, spec, outs[0]), R + C)
    plan.destroy()


if __name__ == "__main__":
    main()
