"""First GPU contact: parity vs numpy on small grids + stage timings at 256^3 / 512^3 (1 GPU)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import accfft_b200 as acc

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
comm = acc.Comm(0, 1, (1, 1))

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))

def check(n, single):
    rng = np.random.default_rng(1)
    rdt, cdt = (np.float32, np.complex64) if single else (np.float64, np.complex128)
    x = rng.uniform(-1, 1, n)
    X = np.fft.rfftn(x)
    mk = acc.accfft_plan_dft_3d_r2c_gpuf if single else acc.accfft_plan_dft_3d_r2c_gpu
    p = mk(n, comm)
    xd = torch.from_numpy(x.astype(rdt)).to(dev)
    out = torch.zeros(p.alloc_max // xd.element_size(), dtype=xd.dtype, device=dev)
    p.execute_r2c(xd, out)
    n2c = n[2] // 2 + 1
    got = out.cpu().numpy().view(cdt)[: n[0] * n[1] * n2c].reshape(n[0], n[1], n2c)
    e1 = rel(got, X)
    back = torch.zeros_like(xd)
    p.execute_c2r(out, back)
    e2 = rel(back.cpu().numpy() / x.size, x)
    lap = torch.zeros_like(xd)
    p.laplace(lap, xd)
    k = [np.fft.fftfreq(m, 1.0 / m) for m in n]
    for i in range(3):
        if n[i] % 2 == 0: k[i][n[i] // 2] = 0
    K2 = k[0][:, None, None] ** 2 + k[1][None, :, None] ** 2 + k[2][None, None, :n2c] ** 2
    lap_ref = np.fft.irfftn(-K2 * X, s=n)
    e3 = rel(lap.cpu().numpy(), lap_ref)
    g = [torch.zeros_like(xd) for _ in range(3)]
    p.grad(g[0], g[1], g[2], xd)
    eg = []
    for ax in range(3):
        kk = [k[0][:, None, None], k[1][None, :, None], k[2][None, None, :n2c]][ax]
        eg.append(rel(g[ax].cpu().numpy(), np.fft.irfftn(1j * kk * X, s=n)))
    dv = torch.zeros_like(xd)
    p.divergence(dv, g[0], g[1], g[2])
    e5 = rel(dv.cpu().numpy(), lap_ref)
    p.destroy()
    # c2c
    z = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    mk = acc.accfft_plan_dft_3d_c2c_gpuf if single else acc.accfft_plan_dft_3d_c2c_gpu
    p = mk(n, comm)
    zd = torch.from_numpy(z.astype(cdt)).to(dev)
    zo = torch.zeros_like(zd)
    p.execute_c2c(acc.ACCFFT_FORWARD, zd, zo)
    e6 = rel(zo.cpu().numpy(), np.fft.fftn(z))
    zb = torch.zeros_like(zd)
    p.execute_c2c(acc.ACCFFT_BACKWARD, zo, zb)
    e7 = rel(zb.cpu().numpy() / z.size, z)
    p.destroy()
    print(f"n={n} single={single}: r2c {e1:.2e} c2r {e2:.2e} lap {e3:.2e} grad {eg[0]:.2e} {eg[1]:.2e} {eg[2]:.2e} div {e5:.2e} c2c {e6:.2e} inv {e7:.2e}", flush=True)

for n in [(8, 8, 8), (16, 12, 20), (64, 64, 64), (48, 96, 40), (128, 128, 128)]:
    for single in (False, True):
        check(n, single)

def bench(n, single, c2c=False, iters=10):
    rdt = torch.float32 if single else torch.float64
    if c2c:
        mk = acc.accfft_plan_dft_3d_c2c_gpuf if single else acc.accfft_plan_dft_3d_c2c_gpu
    else:
        mk = acc.accfft_plan_dft_3d_r2c_gpuf if single else acc.accfft_plan_dft_3d_r2c_gpu
    p = mk(n, comm)
    N = n[0] * n[1] * n[2]
    es = 4 if single else 8
    x = torch.rand((2 * N) if c2c else N, dtype=rdt, device=dev)
    out = torch.zeros(p.alloc_max // es, dtype=rdt, device=dev)
    back = torch.zeros_like(x)
    timer = [0.0] * 5
    def fwd(t=None):
        if c2c: p.execute_c2c(acc.ACCFFT_FORWARD, x, out, t)
        else: p.execute_r2c(x, out, t)
    def inv(t=None):
        if c2c: p.execute_c2c(acc.ACCFFT_BACKWARD, out, back, t)
        else: p.execute_c2r(out, back, t)
    for _ in range(3):
        fwd(); inv()
    p.set_blocking(False)
    p.set_stream(torch.cuda.current_stream().cuda_stream)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    tf = ti = 0.0
    for _ in range(iters):
        e0.record(); fwd(); e1.record(); inv(); e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1); ti += e1.elapsed_time(e2)
    tf /= iters; ti /= iters
    R = N * es; C = (N * 2 * es) if c2c else n[0] * n[1] * (n[2] // 2 + 1) * 2 * es
    algo = 6 * C if c2c else R + 5 * C
    gf = 5 * N * np.log2(N) / 1e9
    print(f"bench n={n} single={single} c2c={c2c}: fwd {tf:.3f} ms inv {ti:.3f} ms  round-trip {2*gf/((tf+ti)*1e-3):.0f} GFLOP/s; "
          f"HBM algo fwd {algo/tf/1e6:.0f} GB/s inv {algo/ti/1e6:.0f} GB/s", flush=True)
    p.set_blocking(True)
    # per-stage (each stage synchronised)
    stage = {}
    p.destroy()

bench((256, 256, 256), False)
bench((512, 512, 512), False)
bench((512, 512, 512), True)
bench((256, 256, 256), False, c2c=True)
bench((512, 512, 512), False, c2c=True)
bench((768, 768, 384), False)
