"""GPU parity tests (run with -m gpu on a B200).  Every call goes through the C ABI of libaccfft_b200.so
(via accfft_b200.api, a plain ctypes binding) and is compared with the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's: relative L2 <= 1e-12 (double), <= 1e-5 (single), on forward spectra and
after the forward+inverse round trip; the reference's own acceptance threshold (relative L1 < 1e-10,
steps/step1/step1_gpu.cpp:103, steps/step5/step5_gpu.cpp:599) is applied to its analytic test functions.
"""
import glob
import os
import threading

import numpy as np
import pytest

import accfft_b200 as acc
from oracle import accfft_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {False: 1e-12, True: 1e-5}
HERE = os.path.dirname(os.path.abspath(__file__))


def torch_mod():
    import torch
    return torch


def dtypes(single):
    return (np.float32, np.complex64) if single else (np.float64, np.complex128)


def to_dev(a, dev):
    torch = torch_mod()
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def alloc_out(plan, single, dev):
    torch = torch_mod()
    es = 4 if single else 8
    return torch.zeros(plan.alloc_max // es, dtype=torch.float32 if single else torch.float64, device=dev)


def spectrum(out, n, single, r2c=True):
    n2c = n[2] // 2 + 1 if r2c else n[2]
    return out.cpu().numpy().view(dtypes(single)[1])[: n[0] * n[1] * n2c].reshape(n[0], n[1], n2c)


def make_plan(n, r2c, single, comm=None, **kw):
    comm = comm or acc.Comm(0, 1, (1, 1))
    return acc.Plan(n, comm, r2c=r2c, single=single, **kw)


@pytest.fixture(params=["tuned", "plain", "tma", "wide"])
def stage_kernels(request, monkeypatch):
    """The strided (Y / X) stages have two kernels: the plain one and the TMA-pipelined persistent one.  Run the
    parity tests under plan-time tuning (ACCFFT_MEASURE, the reference's default flag), and with each kernel forced
    (ACCFFT_B200_TMA is read when a plan is created)."""
    monkeypatch.setenv("ACCFFT_B200_TMA", {"tuned": "1", "plain": "0", "tma": "2", "wide": "3"}[request.param])
    return acc.ACCFFT_MEASURE if request.param == "tuned" else acc.ACCFFT_ESTIMATE


SIZES = [(8, 8, 8), (16, 12, 20), (2, 2, 2), (4, 6, 10), (5, 3, 8), (32, 48, 64), (64, 64, 64), (96, 80, 100),
         (128, 128, 128), (192, 160, 200), (24, 40, 768), (400, 20, 32), (12, 1024, 16), (2048, 8, 12),
         # N2 = 2048: the Z kernels run the 1024-point schedule (a different one in single precision, schedules.def);
         # 1024 / 768 / 1536 along Y and X: lengths that have a wide CTA shape in double
         (4, 6, 2048), (1024, 12, 8), (8, 768, 12), (1536, 4, 6),
         # b, c = 3 families: decimal grids and 27-multiples (G up to 200 threads per column)
         (125, 27, 250), (4, 6, 2000), (1000, 8, 12), (12, 1728, 8), (500, 6, 54), (8, 1500, 10), (6, 10, 1728), (216, 432, 108),
         (135, 270, 6), (375, 4, 540), (8, 675, 750), (1080, 6, 4), (4, 1125, 6), (6, 4, 1350), (864, 4, 6), (450, 6, 4),
         # prime radices 7 / 11 / 13, the composite 14, and 4096-long axes (three radix-16 passes, 256 threads per column)
         (7, 14, 22), (49, 13, 26), (28, 44, 98), (448, 6, 56), (6, 1792, 8), (4, 6, 896), (12, 10, 14),
         (4096, 4, 6), (6, 4096, 4), (4, 6, 8192),
         # N2 = 384: the contiguous kernels run 192 points with a warp per pencil (schedules.def: KEEP_C), partly idle passes
         (6, 10, 384), (20, 12, 192)]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("single", [False, True])
def test_r2c_forward_and_c2r_against_oracle(cuda_device, n, single, stage_kernels):
    x = orc.synthetic_real(n, 20240607)
    X = orc.forward_r2c(x)
    p = make_plan(n, True, single, flags=stage_kernels)
    xd = to_dev(x.astype(dtypes(single)[0]), cuda_device)
    out = alloc_out(p, single, cuda_device)
    p.execute_r2c(xd, out)
    assert orc.rel_l2(spectrum(out, n, single), X) <= TOL[single]
    # inverse of the ORACLE's spectrum (forward-independent), unnormalised
    Xd = to_dev(X.astype(dtypes(single)[1]), cuda_device)
    back = torch_mod().zeros_like(xd)
    p.execute_c2r(Xd.view(xd.dtype), back)
    assert orc.rel_l2(back.cpu().numpy(), x * x.size) <= TOL[single]
    # c2r must not clobber its input here (deliberate fix of src/accfft_gpu.cpp:544-545)
    assert np.array_equal(Xd.cpu().numpy(), X.astype(dtypes(single)[1]))
    p.destroy()


GENERIC_SIZES = [(8, 8, 9), (17, 6, 10), (6, 19, 8), (4, 6, 46), (5, 7, 23), (31, 29, 37), (12, 34, 81), (3, 1021, 4), (6, 5, 7)]


@pytest.mark.parametrize("n", GENERIC_SIZES)
@pytest.mark.parametrize("single", [False, True])
def test_any_length_plans_generic_kernel(cuda_device, n, single):
    """lengths without a compiled schedule (primes >= 17, 81, ...) and odd N2 in r2c plans: the generic O(N^2) stage kernel
    (csrc/generic_dft.cu) -- the reference accepts any length (cuFFT, src/accfft_gpu.cpp:173-253)"""
    torch = torch_mod()
    rdt, cdt = dtypes(single)
    x = orc.synthetic_real(n, 3)
    p = make_plan(n, True, single)
    xd = to_dev(x.astype(rdt), cuda_device)
    out = alloc_out(p, single, cuda_device)
    p.execute_r2c(xd, out)
    assert orc.rel_l2(spectrum(out, n, single), orc.forward_r2c(x)) <= TOL[single]
    back = torch.zeros_like(xd)
    p.execute_c2r(out, back)
    assert orc.rel_l2(back.cpu().numpy() / x.size, x) <= TOL[single]
    tol = 1e-12 if not single else 2e-5
    la = torch.zeros_like(xd)
    p.laplace(la, xd)
    assert orc.rel_l2(la.cpu().numpy(), orc.laplace(x)) <= tol
    g = [torch.zeros_like(xd) for _ in range(3)]
    p.grad(g[0], g[1], g[2], xd)          # per-axis scheme where the axis has kernels, the 3-D scheme (same operator) elsewhere
    for ax in range(3):
        assert orc.rel_l2(g[ax].cpu().numpy(), orc.grad(x, ax)) <= tol, ax
    dv = torch.zeros_like(xd)
    p.divergence(dv, g[0], g[1], g[2])
    assert orc.rel_l2(dv.cpu().numpy(), orc.divergence(*[orc.grad(x, ax) for ax in range(3)])) <= 10 * tol
    p.destroy()
    z = orc.synthetic_complex(n, 4)
    p = make_plan(n, False, single)
    zd = torch.view_as_real(to_dev(z.astype(cdt), cuda_device)).contiguous()
    out = alloc_out(p, single, cuda_device)
    p.execute_c2c(acc.ACCFFT_FORWARD, zd, out)
    assert orc.rel_l2(spectrum(out, n, single, r2c=False), orc.forward_c2c(z)) <= TOL[single]
    back = torch.zeros_like(zd)
    p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)
    assert orc.rel_l2(back.cpu().numpy().ravel().view(cdt).reshape(n) / z.size, z) <= TOL[single]
    p.destroy()


@pytest.mark.parametrize("n", [(8, 8, 8), (16, 12, 20), (6, 10, 5), (64, 64, 64), (48, 96, 40), (128, 128, 128),
                               (20, 24, 1024)])
@pytest.mark.parametrize("single", [False, True])
def test_c2c_against_oracle(cuda_device, n, single, stage_kernels):
    z = orc.synthetic_complex(n, 20240609)
    Z = orc.forward_c2c(z)
    p = make_plan(n, False, single, flags=stage_kernels)
    cdt = dtypes(single)[1]
    zd = to_dev(z.astype(cdt), cuda_device)
    out = torch_mod().zeros_like(zd)
    p.execute_c2c(acc.ACCFFT_FORWARD, zd, out)
    assert orc.rel_l2(out.cpu().numpy(), Z) <= TOL[single]
    back = torch_mod().zeros_like(zd)
    p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)
    assert orc.rel_l2(back.cpu().numpy(), z * z.size) <= TOL[single]
    # in place (steps/step3 default, #define INPLACE step3_gpu.cpp:13)
    buf = zd.clone()
    p.execute_c2c(acc.ACCFFT_FORWARD, buf, buf)
    assert orc.rel_l2(buf.cpu().numpy(), Z) <= TOL[single]
    p.destroy()


@pytest.mark.parametrize("xyz", [0b100, 0b110, 0b101, 0b011, 0b010, 0b001, 0b000])
@pytest.mark.parametrize("n", [(16, 12, 20), (64, 48, 96)])
def test_partial_axis_transforms(cuda_device, n, xyz):
    """std::bitset<3> xyz of accfft_execute_{r2c,c2r,c2c}_gpu (src/accfft_gpu.cpp:477-539, :975-1029): only the selected
    axes are transformed, the layout is always the output layout."""
    torch = torch_mod()
    z = orc.synthetic_complex(n, 77)
    Z = orc.forward_partial(z, xyz, r2c=False)
    p = make_plan(n, False, False)
    zd = to_dev(z, cuda_device)
    out = torch.zeros_like(zd)
    p.execute_c2c(acc.ACCFFT_FORWARD, zd, out, xyz=xyz)
    assert orc.rel_l2(out.cpu().numpy(), Z) <= 1e-12
    back = torch.zeros_like(zd)
    p.execute_c2c(acc.ACCFFT_BACKWARD, out, back, xyz=xyz)
    assert orc.rel_l2(back.cpu().numpy(), orc.inverse_partial(Z, n, xyz, False)) <= 1e-12
    p.destroy()
    p = make_plan(n, True, False)
    x = orc.synthetic_real(n, 78)
    xd = to_dev(x, cuda_device)
    out = alloc_out(p, False, cuda_device)
    if xyz & 4:
        X = orc.forward_partial(x, xyz, r2c=True)
        p.execute_r2c(xd, out, xyz=xyz)
        assert orc.rel_l2(spectrum(out, n, False), X) <= 1e-12
        back = torch.zeros_like(xd)
        p.execute_c2r(out, back, xyz=xyz)
        assert orc.rel_l2(back.cpu().numpy(), orc.inverse_partial(X, n, xyz, True)) <= 1e-12
    else:
        with pytest.raises(acc.AccfftError):   # the real <-> complex step is the Z stage
            p.execute_r2c(xd, out, xyz=xyz)
    p.destroy()


def test_golden_fixtures(cuda_device):
    files = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(f).startswith("ref_"))
    assert files
    for f in files:
        g = np.load(f)
        n = tuple(int(v) for v in g["n"])
        kind = str(g["kind"])
        if kind == "r2c":
            p = make_plan(n, True, False)
            out = alloc_out(p, False, cuda_device)
            p.execute_r2c(to_dev(g["x"], cuda_device), out)
            assert orc.rel_l2(spectrum(out, n, False), g["X"]) <= 1e-12, f
        elif kind == "c2c":
            p = make_plan(n, False, False)
            zd = to_dev(g["x"], cuda_device)
            out = torch_mod().zeros_like(zd)
            p.execute_c2c(acc.ACCFFT_FORWARD, zd, out)
            assert orc.rel_l2(out.cpu().numpy(), g["X"]) <= 1e-12, f
        else:
            p = make_plan(n, True, False)
            xd = to_dev(g["x"], cuda_device)
            o = [torch_mod().zeros_like(xd) for _ in range(5)]
            p.laplace(o[0], xd)
            p.biharmonic(o[1], xd)
            p.grad(o[2], o[3], o[4], xd)
            for t, name in zip(o, ("laplace", "biharmonic", "gx", "gy", "gz")):
                assert orc.rel_l2(t.cpu().numpy(), g[name]) <= 1e-12, (f, name)
        p.destroy()


@pytest.mark.parametrize("n", [(16, 12, 20), (64, 64, 64), (30, 20, 96)])
def test_inplace_r2c_padded_layout(cuda_device, n):
    # steps/step2: data_out == data, rows padded to 2*(N2/2+1) reals (step2_gpu.cpp:41-48,112,127)
    torch = torch_mod()
    x = orc.synthetic_real(n, 5)
    n2c = n[2] // 2 + 1
    p = make_plan(n, True, False)
    buf = alloc_out(p, False, cuda_device)
    padded = np.zeros((n[0], n[1], 2 * n2c))
    padded[:, :, : n[2]] = x
    buf[: padded.size] = torch.from_numpy(padded.ravel()).to(cuda_device)
    p.execute_r2c(buf, buf)
    assert orc.rel_l2(spectrum(buf, n, False), orc.forward_r2c(x)) <= 1e-12
    p.execute_c2r(buf, buf)
    back = buf.cpu().numpy()[: padded.size].reshape(padded.shape)[:, :, : n[2]]
    assert orc.rel_l2(back, x * x.size) <= 1e-12
    p.destroy()


def test_step1_roundtrip_128_reference_threshold(cuda_device):
    # BASELINE config 1 / steps/step1: 128^3 Gaussian, |c2r(r2c(x))/N - x|_1/|x|_1 < 1e-10
    n = (128, 128, 128)
    x = orc.step1_testcase(n)
    p = make_plan(n, True, False)
    xd = to_dev(x, cuda_device)
    out = alloc_out(p, False, cuda_device)
    back = torch_mod().zeros_like(xd)
    p.execute_r2c(xd, out)
    p.execute_c2r(out, back)
    got = back.cpu().numpy() / x.size
    assert np.abs(got - x).sum() / np.abs(x).sum() < 1e-10
    p.destroy()


def test_step3_c2c_roundtrip_reference_threshold(cuda_device):
    n = (128, 128, 128)
    g = orc.step1_testcase(n)
    z = g + 1j * g
    p = make_plan(n, False, False)
    zd = to_dev(z, cuda_device)
    p.execute_c2c(acc.ACCFFT_FORWARD, zd, zd)
    p.execute_c2c(acc.ACCFFT_BACKWARD, zd, zd)
    got = zd.cpu().numpy() / z.size
    assert np.abs(got - z).sum() / np.abs(z).sum() < 1e-10
    p.destroy()


@pytest.mark.parametrize("n,single", [((16, 20, 24), False), ((64, 64, 64), False), ((48, 40, 36), False),
                                      ((64, 64, 64), True), ((10, 6, 12), False)])
def test_operators_against_oracle(cuda_device, n, single):
    torch = torch_mod()
    x = orc.synthetic_real(n, 77)
    p = make_plan(n, True, single)
    xd = to_dev(x.astype(dtypes(single)[0]), cuda_device)
    tol = 1e-12 if not single else 2e-5
    la, bi = torch.zeros_like(xd), torch.zeros_like(xd)
    p.laplace(la, xd)
    p.biharmonic(bi, xd)
    assert orc.rel_l2(la.cpu().numpy(), orc.laplace(x)) <= tol
    assert orc.rel_l2(bi.cpu().numpy(), orc.biharmonic(x)) <= tol
    g = [torch.zeros_like(xd) for _ in range(3)]
    p.grad(g[0], g[1], g[2], xd)
    for ax in range(3):
        assert orc.rel_l2(g[ax].cpu().numpy(), orc.grad(x, ax)) <= tol
    # component selection (std::bitset<3> XYZ of accfft_grad_gpu): only y
    gy = torch.full_like(xd, 123.0)
    p.grad(None, gy, None, xd, xyz=0b010)
    assert orc.rel_l2(gy.cpu().numpy(), orc.grad(x, 1)) <= tol
    a = [orc.synthetic_real(n, 80 + i) for i in range(3)]
    ad = [to_dev(v.astype(dtypes(single)[0]), cuda_device) for v in a]
    dv = torch.zeros_like(xd)
    p.divergence(dv, ad[0], ad[1], ad[2])
    assert orc.rel_l2(dv.cpu().numpy(), orc.divergence(*a)) <= tol
    p.destroy()


@pytest.mark.parametrize("n,single", [((16, 20, 24), False), ((64, 64, 64), False), ((48, 40, 36), False), ((9, 15, 10), False),
                                      ((30, 6, 50), False), ((64, 64, 64), True), ((128, 96, 160), False), ((20, 18, 12), True),
                                      ((1024, 4, 6), True), ((6, 1024, 8), True), ((4, 6, 1024), True), ((768, 6, 4), False),
                                      ((1000, 4, 6), False), ((6, 2000, 4), False), ((4, 6, 1000), True), ((125, 27, 54), False)])
def test_single_axis_transforms_and_slow_operators(cuda_device, n, single):
    """accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t (src/accfft_gpu.cpp:1212-1604) against numpy's rfft along one axis, and the
    reference's two operator schemes (per-axis = default, full 3-D = *_slow) against the oracle and each other."""
    torch = torch_mod()
    rdt, cdt = dtypes(single)
    es = 4 if single else 8
    tol = 1e-12 if not single else 2e-5
    x = orc.synthetic_real(n, 91)
    p = make_plan(n, True, single)
    xd = to_dev(x.astype(rdt), cuda_device)
    for axis in (0, 1, 2):
        Xa = orc.axis_forward(x, axis)
        size, _ = p.axis_size(axis)
        if axis == 0:
            assert size[0] == n[0] // 2 + 1 and size[1] == n[1] * n[2]
        out = torch.zeros(p.alloc_max // es, dtype=xd.dtype, device=cuda_device)
        p.execute_r2c_axis(axis, xd, out)
        got = out.cpu().numpy().view(cdt)[:Xa.size].reshape(Xa.shape)
        assert orc.rel_l2(got, Xa) <= tol, axis
        back = torch.zeros_like(xd)
        p.execute_c2r_axis("xyz"[axis], out, back)
        assert orc.rel_l2(back.cpu().numpy() / n[axis], x) <= tol, axis
        # spectrum untouched by the inverse (the reference's c2r_x transforms its input in place)
        assert np.array_equal(out.cpu().numpy().view(cdt)[:Xa.size].reshape(Xa.shape), got)
    g = [torch.zeros_like(xd) for _ in range(3)]
    gs = [torch.zeros_like(xd) for _ in range(3)]
    p.grad(g[0], g[1], g[2], xd)
    p.grad_slow(gs[0], gs[1], gs[2], xd)
    for ax in range(3):
        assert orc.rel_l2(g[ax].cpu().numpy(), orc.grad(x, ax)) <= tol
        assert orc.rel_l2(gs[ax].cpu().numpy(), orc.grad_slow(x, ax)) <= tol   # differs from grad() for odd N (see oracle)
    a = [orc.synthetic_real(n, 95 + i) for i in range(3)]
    ad = [to_dev(v.astype(rdt), cuda_device) for v in a]
    dv, dvs = torch.full_like(xd, 7.0), torch.full_like(xd, 7.0)
    p.divergence(dv, ad[0], ad[1], ad[2])
    p.divergence_slow(dvs, ad[0], ad[1], ad[2])
    assert orc.rel_l2(dv.cpu().numpy(), orc.divergence(*a)) <= tol
    assert orc.rel_l2(dvs.cpu().numpy(), orc.divergence_slow(*a)) <= tol
    alias = ad[1].clone()
    p.divergence(alias, ad[0], alias, ad[2])         # output aliasing an input component: served by the 3-D scheme, same operator
    assert orc.rel_l2(alias.cpu().numpy(), orc.divergence(*a)) <= tol
    with pytest.raises(acc.AccfftError):
        p.execute_r2c_axis(3, xd, dv)
    p.destroy()


@pytest.mark.parametrize("n,single", [((16, 20, 24), False), ((64, 64, 64), False), ((48, 40, 36), False), ((64, 64, 64), True)])
def test_solver_symbols_against_oracle(cuda_device, n, single):
    """inv_laplace / inv_biharmonic (src/operators.txx:275-392, CPU-only in the reference), Poisson filter (steps/step4/kernels.cu)
    and heat steps (steps/step6) as fused forward -> symbol -> inverse calls"""
    torch = torch_mod()
    rdt, _ = dtypes(single)
    tol = 1e-12 if not single else 2e-5
    x = orc.synthetic_real(n, 131)
    p = make_plan(n, True, single)
    xd = to_dev(x.astype(rdt), cuda_device)
    o = torch.zeros_like(xd)
    p.inv_laplace(o, xd)
    assert orc.rel_l2(o.cpu().numpy(), orc.inv_laplace(x)) <= tol
    p.inv_biharmonic(o, xd)
    assert orc.rel_l2(o.cpu().numpy(), orc.inv_biharmonic(x)) <= tol
    dt = 1.0 / n[0] / n[0] / 4      # steps/step6/step6.cpp:111
    p.symbol_solve(o, xd, 1.0, -dt, False)
    assert orc.rel_l2(o.cpu().numpy(), x + dt * orc.laplace(x)) <= tol
    p.symbol_solve(o, xd, 1.0, dt, True)
    assert orc.rel_l2(o.cpu().numpy(), orc.symbol_solve(x, 1.0, dt, True)) <= tol
    p.symbol_solve(o, xd, 0.0, -4 * np.pi ** 2, True)
    assert orc.rel_l2(o.cpu().numpy(), orc.symbol_solve(x, 0.0, -4 * np.pi ** 2, True)) <= tol
    # Poisson round trip: laplace(inv_laplace(f)) == f for a zero-mean field without Nyquist content
    f = orc.laplace(x)
    fd = to_dev(f.astype(rdt), cuda_device)
    u, lu = torch.zeros_like(xd), torch.zeros_like(xd)
    p.inv_laplace(u, fd)
    p.laplace(lu, u)
    assert orc.rel_l2(lu.cpu().numpy(), f) <= (1e-11 if not single else 1e-3)
    p.destroy()


def test_full_size_per_axis_gradient_512(cuda_device):
    """BASELINE size: the fused per-axis derivative kernels against torch.fft (test-only cross-check on the device)."""
    torch = torch_mod()
    n = (512, 512, 512)
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    xd = torch.rand(n, dtype=torch.float64, device=cuda_device, generator=gen) * 2 - 1
    p = make_plan(n, True, False)
    g = [torch.zeros_like(xd) for _ in range(3)]
    p.grad(g[0], g[1], g[2], xd)
    for ax in range(3):
        k = torch.fft.fftfreq(n[ax], d=1.0 / n[ax], device=cuda_device, dtype=torch.float64)
        k[n[ax] // 2] = 0
        shape = [1, 1, 1]
        shape[ax] = n[ax]
        ref = torch.fft.ifft(torch.fft.fft(xd, dim=ax) * (1j * k.reshape(shape)), dim=ax).real
        err = float(torch.linalg.vector_norm(g[ax] - ref) / torch.linalg.vector_norm(ref))
        assert err <= 1e-12, (ax, err)
        del ref
    p.destroy()


def test_step5_operators_closed_form_256(cuda_device):
    # steps/step5/step5_gpu.cpp: 256^3, f = sin x sin y sin z + exp(-32 |x-pi|^2); relative L1 < 1e-10 (:599)
    torch = torch_mod()
    n = (256, 256, 256)
    t = orc.step5_testcase(n)
    p = make_plan(n, True, False)
    fd = to_dev(t["f"], cuda_device)
    o = [torch.zeros_like(fd) for _ in range(6)]
    p.grad(o[0], o[1], o[2], fd)
    p.laplace(o[3], fd)
    p.biharmonic(o[4], fd)
    p.divergence(o[5], o[0], o[1], o[2])

    def l1(a, b):
        return float(np.abs(a - b).sum() / np.abs(b).sum())

    assert l1(o[0].cpu().numpy(), t["gx"]) < 1e-10
    assert l1(o[1].cpu().numpy(), t["gy"]) < 1e-10
    assert l1(o[2].cpu().numpy(), t["gz"]) < 1e-10
    assert l1(o[3].cpu().numpy(), t["laplace"]) < 1e-10
    # |k|^4 amplifies rounding: the float64 oracle itself is 3.9e-9 away from the closed form at 256^3
    assert l1(o[4].cpu().numpy(), t["biharmonic"]) < 2e-8
    # two correct double-precision FFTs differ by rounding noise that |k|^4 (up to 2.4e9 here) amplifies
    assert orc.rel_l2(o[4].cpu().numpy(), orc.biharmonic(t["f"])) < 5e-9
    assert l1(o[5].cpu().numpy(), t["laplace"]) < 1e-10
    p.destroy()


@pytest.mark.parametrize("n,single", [((512, 512, 512), False), ((256, 256, 256), False), ((768, 768, 384), False),
                                      ((512, 512, 512), True)])
def test_full_size_properties(cuda_device, n, single):
    """BASELINE sizes, size-independent properties (the oracle is too slow / too big to ship vectors for):
    round trip, linearity, Parseval, impulse -> plane wave, and a cuFFT cross-check (torch.fft; test only)."""
    torch = torch_mod()
    rdt = torch.float32 if single else torch.float64
    tol = TOL[single]
    gen = torch.Generator(device=cuda_device).manual_seed(1234)
    N = n[0] * n[1] * n[2]
    n2c = n[2] // 2 + 1
    p = make_plan(n, True, single)
    x = torch.rand(n, dtype=rdt, device=cuda_device, generator=gen) * 2 - 1
    out = alloc_out(p, single, cuda_device)
    p.execute_r2c(x, out)
    X = torch.view_as_complex(out[: 2 * n[0] * n[1] * n2c].view(n[0], n[1], n2c, 2))
    # cuFFT cross-check
    ref = torch.fft.rfftn(x)
    assert float(torch.linalg.vector_norm(X - ref) / torch.linalg.vector_norm(ref)) <= tol
    # Parseval with Hermitian weights
    w = torch.full((n2c,), 2.0, dtype=rdt, device=cuda_device)
    w[0] = 1.0
    if n[2] % 2 == 0:
        w[-1] = 1.0
    e_spec = float(((X.real ** 2 + X.imag ** 2) * w).sum()) / N
    e_phys = float((x * x).sum())
    assert abs(e_spec - e_phys) <= (1e-10 if not single else 1e-4) * e_phys
    # round trip
    back = torch.zeros_like(x)
    p.execute_c2r(out, back)
    assert float(torch.linalg.vector_norm(back / N - x) / torch.linalg.vector_norm(x)) <= tol
    del ref, back
    # linearity: F(a x + b y) = a F(x) + b F(y)
    y = torch.rand(n, dtype=rdt, device=cuda_device, generator=gen) - 0.5
    out2 = alloc_out(p, single, cuda_device)
    p.execute_r2c(y, out2)
    comb = alloc_out(p, single, cuda_device)
    p.execute_r2c(0.75 * x - 1.5 * y, comb)
    lin = 0.75 * out - 1.5 * out2
    assert float(torch.linalg.vector_norm(comb - lin) / torch.linalg.vector_norm(lin)) <= tol * 2
    del y, out2, comb, lin
    # impulse at (1, 2, 3) -> X[k] = exp(-2 pi i (k0/N0 + 2 k1/N1 + 3 k2/N2))
    x.zero_()
    x[1, 2, 3] = 1.0
    p.execute_r2c(x, out)
    k0 = torch.arange(n[0], device=cuda_device, dtype=torch.float64)[:, None, None] / n[0]
    k1 = torch.arange(n[1], device=cuda_device, dtype=torch.float64)[None, :, None] * 2 / n[1]
    k2 = torch.arange(n2c, device=cuda_device, dtype=torch.float64)[None, None, :] * 3 / n[2]
    ph = -2 * np.pi * (k0 + k1 + k2)
    err = torch.maximum((X.real - torch.cos(ph)).abs().max(), (X.imag - torch.sin(ph)).abs().max())
    assert float(err) <= (1e-13 if not single else 1e-5)
    p.destroy()


def test_full_size_c2c_roundtrip_512(cuda_device):
    torch = torch_mod()
    n = (512, 512, 512)
    p = make_plan(n, False, False)
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    z = torch.view_as_complex(torch.rand(n + (2,), dtype=torch.float64, device=cuda_device, generator=gen) - 0.5)
    out = torch.zeros_like(z)
    p.execute_c2c(acc.ACCFFT_FORWARD, z, out)
    ref = torch.fft.fftn(z)
    assert float(torch.linalg.vector_norm(out - ref) / torch.linalg.vector_norm(ref)) <= 1e-12
    del ref
    p.execute_c2c(acc.ACCFFT_BACKWARD, out, out)
    assert float(torch.linalg.vector_norm(out / z.numel() - z) / torch.linalg.vector_norm(z)) <= 1e-12
    p.destroy()


def test_error_behaviour(cuda_device):
    torch = torch_mod()
    comm = acc.Comm(0, 1, (1, 1))
    with pytest.raises(acc.AccfftError):     # beyond what even the generic stage kernel can stage
        acc.accfft_plan_dft_3d_c2c_gpu((9000, 8, 8), comm)
    q = acc.accfft_plan_dft_3d_r2c_gpu((8, 17, 8), comm)     # plans (generic kernel along Y) ...
    xq = torch.zeros(8 * 17 * 8, dtype=torch.float64, device=cuda_device)
    with pytest.raises(acc.AccfftError):     # ... but the single-axis executes exist for scheduled lengths only
        q.execute_r2c_axis(1, xq, alloc_out(q, False, cuda_device))
    q.destroy()
    p = acc.accfft_plan_dft_3d_r2c_gpu((8, 8, 8), comm)
    x = torch.zeros(8 * 8 * 8, dtype=torch.float64, device=cuda_device)
    out = alloc_out(p, False, cuda_device)
    with pytest.raises(acc.AccfftError):     # wrong plan kind (the reference prints and returns, :641-646)
        p.execute_c2c(acc.ACCFFT_FORWARD, x, out)
    with pytest.raises(acc.AccfftError):     # r2c without the Z axis is not a transform (see accfft_b200.h)
        p.execute_r2c(x, out, xyz=0b011)
    with pytest.raises(acc.AccfftError):     # not a 3-bit mask
        p.execute_r2c(x, out, xyz=8)
    with pytest.raises(acc.AccfftError):
        p.execute_r2c(None, out)
    with pytest.raises(acc.AccfftError):     # host tensor
        p.execute_r2c(torch.zeros(512, dtype=torch.float64), out)
    p.execute_r2c(x, out)                    # still usable afterwards
    p.destroy()
    p.destroy()                              # idempotent
    # a multi-rank plan cannot run before its peers are attached
    q = acc.Plan((8, 8, 8), acc.Comm(0, 2, (1, 2)), r2c=True, single=False, wire=False)
    with pytest.raises(acc.AccfftError):
        q.execute_r2c(x, out)
    acc.lib().accfftb_plan_destroy(q._h)
    q._h = None


def test_barrier_timeout_is_sticky_and_visible_without_blocking(cuda_device, monkeypatch):
    """A rank whose peers never arrive: the barrier kernel times out, the status word is sticky, and every later call fails
    with ACCFFTB_ERR_TIMEOUT -- in non-blocking mode too (accfftb_plan_check / the next call), not only after a blocking
    synchronise (round-1 finding: non-blocking plans silently ran on stale peer data)."""
    torch = torch_mod()
    monkeypatch.setenv("ACCFFT_B200_BARRIER_TIMEOUT_MS", "100")
    n = (16, 12, 20)
    comms = [acc.Comm(r, 2, (2, 1)) for r in range(2)]
    plans = [acc.Plan(n, comms[r], r2c=True, single=False, wire=False) for r in range(2)]
    acc.attach_local_peers(plans)
    p = plans[0]
    x = torch.zeros(p.isize[0] * p.isize[1] * p.isize[2], dtype=torch.float64, device=cuda_device)
    out = alloc_out(p, False, cuda_device)
    p.check()                                   # healthy
    p.set_blocking(False)
    p.execute_r2c(x, out)                       # rank 1 never calls: rank 0's barriers time out on the device
    torch.cuda.synchronize()
    with pytest.raises(acc.AccfftError, match="timed out"):
        p.check()
    with pytest.raises(acc.AccfftError, match="timed out"):
        p.execute_r2c(x, out)                   # refused before anything is launched
    p.set_blocking(True)
    with pytest.raises(acc.AccfftError, match="timed out"):
        p.execute_c2r(out, x)
    for q in plans:
        acc.lib().accfftb_plan_destroy(q._h)
        q._h = None


@pytest.mark.parametrize("n,dims,r2c", [((16, 12, 20), (2, 1), True), ((64, 48, 40), (4, 1), True), ((5, 6, 6), (4, 1), True),
                                        ((32, 24, 20), (8, 1), False), ((128, 96, 80), (2, 1), True)])
def test_copy_engine_exchange_matches_peer_stores(cuda_device, monkeypatch, n, dims, r2c):
    """slab grids: the staged exchange (ACCFFT_B200_CE=<chunks>: staging buffer + copy-engine / TMA-push transfers) and the
    default direct peer-store exchange compute the same butterflies in the same order; only the kernel variant of the
    packing stage may differ (FMA contraction), so spectra and round trips agree to a few ulp"""
    torch = torch_mod()
    nprocs = dims[0] * dims[1]
    x = orc.synthetic_real(n, 41) if r2c else orc.synthetic_complex(n, 41)
    results = {}
    for mode, env in (("ce", {"ACCFFT_B200_CE": "3", "ACCFFT_B200_CE_LANES": "2", "ACCFFT_B200_PUSH": "0"}),
                      ("push", {"ACCFFT_B200_CE": "2", "ACCFFT_B200_CE_LANES": "3", "ACCFFT_B200_PUSH": "8"}), ("direct", {"ACCFFT_B200_CE": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
        plans = [acc.Plan(n, comms[r], r2c=r2c, single=False, wire=False, flags=acc.ACCFFT_ESTIMATE) for r in range(nprocs)]
        acc.attach_local_peers(plans)
        assert plans[0].exchange_mode()["mode"] == ("peer_store" if mode == "direct" else "copy_engine")
        streams = [torch.cuda.Stream() for _ in range(nprocs)]
        ins, outs, backs = [], [], []
        for r in range(nprocs):
            loc = orc.local_input(x, n, dims, comms[r].coords)
            t = to_dev(loc, cuda_device) if loc.size else torch.zeros(2, dtype=torch.float64, device=cuda_device)
            ins.append(torch.view_as_real(t).contiguous() if (not r2c and loc.size) else t)
            outs.append(torch.zeros(max(plans[r].alloc_max // 8, 2), dtype=torch.float64, device=cuda_device))
            backs.append(torch.zeros_like(ins[-1]))
            plans[r].set_stream(streams[r].cuda_stream)
        torch.cuda.synchronize()

        def body(r):
            with torch.cuda.stream(streams[r]):
                for _ in range(3):
                    if r2c:
                        plans[r].execute_r2c(ins[r], outs[r])
                        plans[r].execute_c2r(outs[r], backs[r])
                    else:
                        plans[r].execute_c2c(acc.ACCFFT_FORWARD, ins[r], outs[r])
                        plans[r].execute_c2c(acc.ACCFFT_BACKWARD, outs[r], backs[r])

        run_ranks(nprocs, body)
        torch.cuda.synchronize()
        if mode != "direct":
            assert plans[0].copy_count > 0
        results[mode] = ([o.cpu().numpy().copy() for o in outs], [b.cpu().numpy().copy() for b in backs])
        blocks = [outs[r].cpu().numpy().view(np.complex128)[:int(np.prod(plans[r].osize))].reshape(plans[r].osize) for r in range(nprocs)]
        X = orc.forward_r2c(x) if r2c else orc.forward_c2c(x)
        assert orc.rel_l2(orc.assemble_output(blocks, n, dims, r2c=r2c), X) <= 1e-12
        for p in plans:
            acc.lib().accfftb_plan_destroy(p._h)
            p._h = None
    for a, b in zip(results["ce"][0] + results["ce"][1] + results["push"][0] + results["push"][1],
                    2 * (results["direct"][0] + results["direct"][1])):
        assert orc.rel_l2(a, b) <= 1e-15 if np.any(b) else not np.any(a)


def test_timer_accumulates_and_launch_count(cuda_device):
    torch = torch_mod()
    n = (64, 64, 64)
    p = make_plan(n, True, False)
    x = torch.rand(n, dtype=torch.float64, device=cuda_device)
    out = alloc_out(p, False, cuda_device)
    timer = [0.0] * 5
    c0 = p.launch_count
    p.execute_r2c(x, out, timer)
    t1 = timer[4]
    assert t1 > 0 and timer[0] == 0 and timer[1] == 0 and timer[3] == 0
    p.execute_r2c(x, out, timer)
    assert timer[4] > t1                     # += semantics (doc/timing.md)
    assert p.launch_count - c0 == 6          # 3 stage kernels per transform, no exchange on one rank
    p.set_profiling(True)
    p.execute_r2c(x, out)
    back = torch.zeros_like(x)
    p.execute_c2r(out, back)
    prof = p.profile_read()
    assert p.fused_info(False) is None and p.fused_info(True) is None   # the fused local pair is opt-in (ACCFFT_B200_FUSED=1)
    assert all(prof[k][1] == 1 and prof[k][0] > 0 for k in ("fwd_z", "fwd_y", "fwd_x", "inv_x", "inv_y", "inv_z"))
    assert prof["barrier"][1] == 0
    p.destroy()


# ---- multi-rank path on ONE device: every rank is a host thread with its own stream --------------------
def run_ranks(nprocs, fn):
    errs = []

    def wrap(r):
        try:
            torch_mod().cuda.set_device(0)
            fn(r)
        except Exception as e:  # pragma: no cover
            errs.append((r, repr(e)))

    th = [threading.Thread(target=wrap, args=(r,)) for r in range(nprocs)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not errs, errs
    assert not any(t.is_alive() for t in th), "rank thread hung"


MR_CASES = [((16, 12, 20), (1, 2), True), ((16, 12, 20), (2, 1), True), ((16, 12, 20), (2, 2), True),
            ((12, 10, 16), (2, 2), True), ((5, 6, 6), (2, 2), True), ((64, 48, 40), (2, 2), True),
            ((3, 5, 4), (4, 1), True), ((16, 12, 20), (2, 2), False), ((6, 10, 5), (1, 4), False),
            ((128, 128, 128), (2, 2), True)]
# single precision: blocks of 5,5,5,2 / 9,9,9,6 spectral columns -- odd and even row strides in one process grid, i.e.
# ranks that can and ranks that cannot use the TMA stage kernel side by side, and unaligned column shifts
MR_CASES_F = [((8, 32, 32), (1, 4), True), ((8, 32, 32), (2, 4), True), ((16, 64, 64), (2, 4), True),
              ((32, 32, 12), (2, 2), False), ((64, 64, 64), (2, 2), True)]


@pytest.mark.parametrize("n,dims,r2c,single", [c + (False,) for c in MR_CASES] + [c + (True,) for c in MR_CASES_F])
def test_multirank_threads_on_one_gpu(cuda_device, n, dims, r2c, single, stage_kernels):
    """The exchange path (peer-addressed stores fused into the stage kernels + the inter-GPU barrier kernel)
    with all ranks on cuda:0: uneven and empty pencils included.  Replaces what the reference can only test
    by running under mpirun on a real machine."""
    torch = torch_mod()
    nprocs = dims[0] * dims[1]
    rdt, cdt = dtypes(single)
    tdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    x = orc.synthetic_real(n, 31) if r2c else orc.synthetic_complex(n, 31)
    X = orc.forward_r2c(x) if r2c else orc.forward_c2c(x)
    comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
    plans = [acc.Plan(n, comms[r], r2c=r2c, single=single, wire=False, flags=stage_kernels) for r in range(nprocs)]
    acc.attach_local_peers(plans)
    streams = [torch.cuda.Stream() for _ in range(nprocs)]
    ins, outs, backs = [], [], []
    for r in range(nprocs):
        loc = orc.local_input(x, n, dims, comms[r].coords).astype(rdt if r2c else cdt)
        t = to_dev(loc, cuda_device) if loc.size else torch.zeros(2, dtype=tdt, device=cuda_device)
        ins.append(torch.view_as_real(t).contiguous() if (not r2c and loc.size) else t)
        outs.append(torch.zeros(max(plans[r].alloc_max // es, 2), dtype=tdt, device=cuda_device))
        backs.append(torch.zeros_like(ins[-1]))
        plans[r].set_stream(streams[r].cuda_stream)
    torch.cuda.synchronize()

    def body(r):
        with torch.cuda.stream(streams[r]):
            for _ in range(3):   # repeated calls exercise buffer-reuse ordering across ranks
                if r2c:
                    plans[r].execute_r2c(ins[r], outs[r])
                    plans[r].execute_c2r(outs[r], backs[r])
                else:
                    plans[r].execute_c2c(acc.ACCFFT_FORWARD, ins[r], outs[r])
                    plans[r].execute_c2c(acc.ACCFFT_BACKWARD, outs[r], backs[r])

    run_ranks(nprocs, body)
    torch.cuda.synchronize()
    blocks = []
    for r in range(nprocs):
        osz = plans[r].osize
        cnt = osz[0] * osz[1] * osz[2]
        blocks.append(outs[r].cpu().numpy().view(cdt)[:cnt].reshape(osz))
    assert orc.rel_l2(orc.assemble_output(blocks, n, dims, r2c=r2c), X) <= TOL[single]
    bl = []
    for r in range(nprocs):
        isz = plans[r].isize
        cnt = isz[0] * isz[1] * isz[2]
        b = backs[r].cpu().numpy().ravel()
        b = b.view(cdt) if not r2c else b
        bl.append(b[:cnt].reshape(isz))
    assert orc.rel_l2(orc.assemble_input(bl, n, dims), x * x.size) <= TOL[single]
    for p in plans:
        acc.lib().accfftb_plan_destroy(p._h)
        p._h = None


@pytest.mark.parametrize("n,dims,single", [((16, 12, 20), (2, 2), False), ((12, 10, 16), (2, 4), False), ((5, 6, 6), (4, 2), False),
                                           ((64, 48, 40), (1, 2), False), ((64, 48, 40), (2, 1), False),
                                           ((16, 64, 64), (2, 4), True), ((128, 96, 160), (2, 4), False)])
def test_multirank_single_axis_threads(cuda_device, n, dims, single):
    """single-axis real paths with their real-data exchanges (T_plan_y / T_plan_x of the reference,
    src/accfft_gpu.cpp:369-373, :426-438) on one device, every rank a host thread"""
    torch = torch_mod()
    nprocs = dims[0] * dims[1]
    rdt, cdt = dtypes(single)
    tdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    tol = TOL[single]
    x = orc.synthetic_real(n, 51)
    comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
    plans = [acc.Plan(n, comms[r], r2c=True, single=single, wire=False) for r in range(nprocs)]
    acc.attach_local_peers(plans)
    streams = [torch.cuda.Stream() for _ in range(nprocs)]
    ins, spec, backs = [], [], []
    for r in range(nprocs):
        loc = orc.local_input(x, n, dims, comms[r].coords).astype(rdt)
        ins.append(to_dev(loc, cuda_device) if loc.size else torch.zeros(2, dtype=tdt, device=cuda_device))
        spec.append([torch.zeros(max(plans[r].alloc_max // es, 2), dtype=tdt, device=cuda_device) for _ in range(3)])
        backs.append([torch.zeros_like(ins[-1]) for _ in range(3)])
        plans[r].set_stream(streams[r].cuda_stream)
    torch.cuda.synchronize()

    def body(r):
        with torch.cuda.stream(streams[r]):
            for _ in range(2):   # repeated: exchange-buffer reuse across operations
                for axis in (0, 1, 2):
                    plans[r].execute_r2c_axis(axis, ins[r], spec[r][axis])
                    plans[r].execute_c2r_axis(axis, spec[r][axis], backs[r][axis])

    run_ranks(nprocs, body)
    torch.cuda.synchronize()
    for axis in (0, 1, 2):
        Xa = orc.axis_forward(x, axis)
        num = den = 0.0
        for r in range(nprocs):
            ref = orc.axis_local_spectrum(Xa, n, dims, comms[r].coords, axis)
            size, _ = plans[r].axis_size(axis)
            assert int(np.prod(size)) == ref.size
            got = spec[r][axis].cpu().numpy().view(cdt)[:ref.size].reshape(ref.shape)
            num += float(np.sum(np.abs(got - ref) ** 2))
            den += float(np.sum(np.abs(ref) ** 2))
        assert np.sqrt(num / den) <= tol, axis
        bl = []
        for r in range(nprocs):
            isz = plans[r].isize
            bl.append(backs[r][axis].cpu().numpy().ravel()[:isz[0] * isz[1] * isz[2]].reshape(isz))
        assert orc.rel_l2(orc.assemble_input(bl, n, dims), x * n[axis]) <= tol, axis
    for p in plans:
        acc.lib().accfftb_plan_destroy(p._h)
        p._h = None


def test_multirank_laplace_threads(cuda_device):
    torch = torch_mod()
    n, dims = (16, 20, 24), (2, 2)
    nprocs = 4
    x = orc.synthetic_real(n, 41)
    comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
    plans = [acc.Plan(n, comms[r], r2c=True, single=False, wire=False) for r in range(nprocs)]
    acc.attach_local_peers(plans)
    streams = [torch.cuda.Stream() for _ in range(nprocs)]
    ins = [to_dev(orc.local_input(x, n, dims, comms[r].coords), cuda_device) for r in range(nprocs)]
    res = {k: [torch.zeros_like(t) for t in ins] for k in ("lap", "gx", "gy", "gz", "div", "sx", "sy", "sz", "sdiv")}
    for r in range(nprocs):
        plans[r].set_stream(streams[r].cuda_stream)
    torch.cuda.synchronize()

    def body(r):
        with torch.cuda.stream(streams[r]):
            plans[r].laplace(res["lap"][r], ins[r])
            plans[r].grad(res["gx"][r], res["gy"][r], res["gz"][r], ins[r])
            plans[r].divergence(res["div"][r], res["gx"][r], res["gy"][r], res["gz"][r])
            plans[r].grad_slow(res["sx"][r], res["sy"][r], res["sz"][r], ins[r])
            plans[r].divergence_slow(res["sdiv"][r], res["sx"][r], res["sy"][r], res["sz"][r])

    run_ranks(nprocs, body)
    torch.cuda.synchronize()
    asm = lambda k: orc.assemble_input([t.cpu().numpy() for t in res[k]], n, dims)
    assert orc.rel_l2(asm("lap"), orc.laplace(x)) <= 1e-12
    for ax, k in enumerate(("gx", "gy", "gz")):
        assert orc.rel_l2(asm(k), orc.grad(x, ax)) <= 1e-12
        assert orc.rel_l2(asm("s" + k[1]), orc.grad(x, ax)) <= 1e-12
    assert orc.rel_l2(asm("div"), orc.divergence(orc.grad(x, 0), orc.grad(x, 1), orc.grad(x, 2))) <= 1e-12
    assert orc.rel_l2(asm("sdiv"), orc.divergence(orc.grad(x, 0), orc.grad(x, 1), orc.grad(x, 2))) <= 1e-12
    for p in plans:
        acc.lib().accfftb_plan_destroy(p._h)
        p._h = None


@pytest.mark.parametrize("n,dims", [((16, 12, 20), (2, 1)), ((64, 48, 40), (4, 1)), ((12, 10, 16), (2, 2)), ((5, 6, 6), (1, 4))])
def test_double_buffered_exchange_mixed_operations(cuda_device, monkeypatch, n, dims):
    """Double-buffered exchange buffers (default for multi-rank plans): one barrier per transform on slab grids.  A long,
    mixed sequence of operations per rank, back to back and without host synchronisation between ranks, must give the same
    results as with single buffering (ACCFFT_B200_DB=0, every reuse barrier in place) and match the oracle."""
    torch = torch_mod()
    nprocs = dims[0] * dims[1]
    x = orc.synthetic_real(n, 51)
    X = orc.forward_r2c(x)
    outs_by_mode = {}
    for mode in ("db", "single"):
        monkeypatch.setenv("ACCFFT_B200_DB", "1" if mode == "db" else "0")
        comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
        plans = [acc.Plan(n, comms[r], r2c=True, single=False, wire=False, flags=acc.ACCFFT_ESTIMATE) for r in range(nprocs)]
        acc.attach_local_peers(plans)
        streams = [torch.cuda.Stream() for _ in range(nprocs)]
        res = [dict() for _ in range(nprocs)]
        bar0 = [plans[r].launch_count for r in range(nprocs)]

        def body(r):
            p = plans[r]
            p.set_stream(streams[r].cuda_stream)
            loc = orc.local_input(x, n, dims, comms[r].coords)
            with torch.cuda.stream(streams[r]):
                xd = to_dev(loc, cuda_device) if loc.size else torch.zeros(2, dtype=torch.float64, device=cuda_device)
                out = torch.zeros(max(p.alloc_max // 8, 2), dtype=torch.float64, device=cuda_device)
                back, la = torch.zeros_like(xd), torch.zeros_like(xd)
                g = [torch.zeros_like(xd) for _ in range(3)]
                p.set_blocking(False)
                for _ in range(3):     # nothing but stream order and the plan's own barriers separates these
                    p.execute_r2c(xd, out)
                    p.execute_c2r(out, back)
                    p.execute_c2r(out, back)
                    p.laplace(la, xd)
                    p.grad(g[0], g[1], g[2], xd)
                    p.execute_r2c(xd, out)
                    p.grad_slow(g[0], g[1], g[2], xd)
                    p.execute_r2c_axis(1, xd, out)
                    p.execute_c2r_axis(1, out, back)
                    p.execute_r2c(xd, out)
                    p.execute_c2r(out, back)
                torch.cuda.synchronize()
                res[r] = {"out": out.cpu().numpy().copy(), "back": back.cpu().numpy().copy(), "la": la.cpu().numpy().copy(),
                          "g": [t.cpu().numpy().copy() for t in g], "loc": loc}

        run_ranks(nprocs, body)
        torch.cuda.synchronize()
        for r in range(nprocs):
            p, loc = plans[r], res[r]["loc"]
            cnt = int(np.prod(p.osize))
            if cnt:
                got = res[r]["out"].view(np.complex128)[:cnt].reshape(p.osize)
                assert orc.rel_l2(got, orc.local_output(X, n, dims, comms[r].coords, r2c=True)) <= 1e-12
            if loc.size:
                assert orc.rel_l2(res[r]["back"].reshape(loc.shape) / x.size, loc) <= 1e-12
                assert orc.rel_l2(res[r]["la"].reshape(loc.shape), orc.local_input(orc.laplace(x), n, dims, comms[r].coords)) <= 1e-12
                for ax in range(3):
                    assert orc.rel_l2(res[r]["g"][ax].reshape(loc.shape),
                                      orc.local_input(orc.grad_slow(x, ax), n, dims, comms[r].coords)) <= 1e-12
        outs_by_mode[mode] = (res, [plans[r].launch_count - bar0[r] for r in range(nprocs)])
        for p in plans:
            acc.lib().accfftb_plan_destroy(p._h)
            p._h = None
    # fewer kernels (barriers) with double buffering on slab grids, never more
    assert all(a <= b for a, b in zip(outs_by_mode["db"][1], outs_by_mode["single"][1]))
    if dims[1] == 1 or dims[0] == 1:
        assert outs_by_mode["db"][1][0] < outs_by_mode["single"][1][0]


# ---- fused local Z / Y pair (csrc/fused.cuh): one persistent kernel, intermediate planes in an L2-resident ring ----------
FUSED_CASES = [((256, 256, 256), True, False), ((64, 64, 64), True, False), ((64, 64, 64), True, True), ((40, 128, 128), True, False), ((24, 256, 256), True, True),
               ((37, 64, 64), True, False), ((64, 64, 64), False, False), ((20, 128, 128), False, True), ((3, 512, 512), True, False)]


@pytest.mark.parametrize("ring", ["auto", "tiny"])
@pytest.mark.parametrize("n,r2c,single", FUSED_CASES)
def test_fused_local_pair_matches_separate_kernels(cuda_device, monkeypatch, n, r2c, single, ring):
    """The fused kernel runs the same tile bodies as the stand-alone Z and Y kernels, so its results must be BITWISE those of
    the two-kernel path -- with the ring spanning the whole buffer ("auto" at these sizes) and with a ring of two planes and the
    second pass one chunk behind the first ("tiny": every tile waits on its dependencies, slots are overwritten constantly)."""
    torch = torch_mod()
    rdt, cdt = dtypes(single)
    x = orc.synthetic_real(n, 77) if r2c else orc.synthetic_complex(n, 77)
    X = orc.forward_r2c(x) if r2c else orc.forward_c2c(x)
    res = {}
    for mode in ("fused", "separate"):
        monkeypatch.setenv("ACCFFT_B200_FUSED", "1" if mode == "fused" else "0")
        if mode == "fused" and ring == "tiny":
            for k, v in (("CP", "1"), ("LAG", "1"), ("NS", "2")):
                monkeypatch.setenv("ACCFFT_B200_FUSED_" + k, v)
        p = make_plan(n, r2c, single, flags=acc.ACCFFT_ESTIMATE)
        xd = to_dev(x.astype(rdt if r2c else cdt), cuda_device)
        xin = xd if r2c else torch.view_as_real(xd).contiguous()
        out = alloc_out(p, single, cuda_device)
        back = torch.zeros_like(xin)
        for _ in range(3):
            if r2c:
                p.execute_r2c(xin, out)
                p.execute_c2r(out, back)
            else:
                p.execute_c2c(acc.ACCFFT_FORWARD, xin, out)
                p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)
        fi, ii = p.fused_info(False), p.fused_info(True)
        if mode == "fused":
            assert fi is not None and ii is not None, "no fused kernel ran for a length pair of fused_combos.def"
            if ring == "tiny" and n[0] > 2:
                assert fi["ring_planes"] == 2 and ii["ring_planes"] == 2 and fi["lag"] == 1
        else:
            assert fi is None and ii is None
        res[mode] = (out.cpu().numpy().copy(), back.cpu().numpy().copy())
        assert orc.rel_l2(spectrum(out, n, single, r2c), X) <= TOL[single]
        b = back.cpu().numpy()
        b = b.view(cdt).reshape(n) if not r2c else b
        assert orc.rel_l2(b, x * x.size) <= TOL[single]
        # operators ride on the same pair (laplace: forward pair + symbol on X + inverse pair)
        if r2c:
            lap = torch.zeros_like(xin)
            p.laplace(lap, xin)
            res[mode] += (lap.cpu().numpy().copy(),)
            assert orc.rel_l2(lap.cpu().numpy(), orc.laplace(x)) <= (1e-11 if not single else 2e-4)
        p.destroy()
        for k in ("CP", "LAG", "NS"):
            monkeypatch.delenv("ACCFFT_B200_FUSED_" + k, raising=False)
    for a, b in zip(res["fused"], res["separate"]):
        if single:   # the stand-alone single-precision Z kernels run their own schedules (schedules.def: KEEP_FC): same accuracy, not the same bits
            assert orc.rel_l2(a, b) <= 2e-6
        else:
            assert np.array_equal(a, b)


@pytest.mark.parametrize("n,dims,r2c,single", [((64, 64, 64), (2, 1), True, False), ((64, 64, 64), (4, 1), True, True),
                                               ((30, 128, 128), (4, 1), True, False), ((64, 64, 64), (2, 1), False, False),
                                               ((16, 64, 64), (8, 1), True, False)])
@pytest.mark.parametrize("ring", ["auto", "tiny", "default"])
def test_fused_local_pair_on_slab_grids_threads(cuda_device, monkeypatch, n, dims, r2c, single, ring):
    """slab grids: forward = fused Z + peer-storing Y (general strided body), inverse = fused local Y + Z behind the exchange;
    ranks as host threads on cuda:0, uneven and empty slabs included (30 planes over 4 ranks, 16 over 8)"""
    torch = torch_mod()
    nprocs = dims[0] * dims[1]
    rdt, cdt = dtypes(single)
    tdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    # "default": no switch -- multi-rank slab plans fuse the forward pair (its Y pass stores to peers) and only that one
    if ring != "default":
        monkeypatch.setenv("ACCFFT_B200_FUSED", "1")
    if ring == "tiny":
        for k, v in (("CP", "1"), ("LAG", "1"), ("NS", "2")):
            monkeypatch.setenv("ACCFFT_B200_FUSED_" + k, v)
    x = orc.synthetic_real(n, 91) if r2c else orc.synthetic_complex(n, 91)
    X = orc.forward_r2c(x) if r2c else orc.forward_c2c(x)
    comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
    plans = [acc.Plan(n, comms[r], r2c=r2c, single=single, wire=False, flags=acc.ACCFFT_ESTIMATE) for r in range(nprocs)]
    acc.attach_local_peers(plans)
    streams = [torch.cuda.Stream() for _ in range(nprocs)]
    ins, outs, backs = [], [], []
    for r in range(nprocs):
        loc = orc.local_input(x, n, dims, comms[r].coords).astype(rdt if r2c else cdt)
        t = to_dev(loc, cuda_device) if loc.size else torch.zeros(2, dtype=tdt, device=cuda_device)
        ins.append(torch.view_as_real(t).contiguous() if (not r2c and loc.size) else t)
        outs.append(torch.zeros(max(plans[r].alloc_max // es, 2), dtype=tdt, device=cuda_device))
        backs.append(torch.zeros_like(ins[-1]))
        plans[r].set_stream(streams[r].cuda_stream)
    torch.cuda.synchronize()

    def body(r):
        with torch.cuda.stream(streams[r]):
            for _ in range(3):
                if r2c:
                    plans[r].execute_r2c(ins[r], outs[r])
                    plans[r].execute_c2r(outs[r], backs[r])
                else:
                    plans[r].execute_c2c(acc.ACCFFT_FORWARD, ins[r], outs[r])
                    plans[r].execute_c2c(acc.ACCFFT_BACKWARD, outs[r], backs[r])

    run_ranks(nprocs, body)
    torch.cuda.synchronize()
    # default mode: r2c plans fuse the forward pair only, c2c plans nothing (measured for r2c only)
    assert (plans[0].fused_info(False) is not None) == (ring != "default" or r2c)
    assert (plans[0].fused_info(True) is not None) == (ring != "default")
    blocks = []
    for r in range(nprocs):
        osz = plans[r].osize
        blocks.append(outs[r].cpu().numpy().view(cdt)[:osz[0] * osz[1] * osz[2]].reshape(osz))
    assert orc.rel_l2(orc.assemble_output(blocks, n, dims, r2c=r2c), X) <= TOL[single]
    bl = []
    for r in range(nprocs):
        isz = plans[r].isize
        b = backs[r].cpu().numpy().ravel()
        b = b.view(cdt) if not r2c else b
        bl.append(b[:isz[0] * isz[1] * isz[2]].reshape(isz))
    assert orc.rel_l2(orc.assemble_input(bl, n, dims), x * x.size) <= TOL[single]
    for p in plans:
        acc.lib().accfftb_plan_destroy(p._h)
        p._h = None


# ---- batched fields (accfftb_plan_create_many / accfftb_execute_*_many) ---------------------------------------------------
@pytest.mark.parametrize("n,dims,r2c,single", [((64, 48, 40), (1, 1), True, False), ((16, 12, 20), (2, 2), True, False),
                                               ((5, 6, 6), (4, 1), True, False), ((64, 48, 40), (2, 2), True, True),
                                               ((16, 12, 20), (2, 2), False, False), ((3, 5, 4), (4, 2), True, False)])
def test_batched_fields_equal_single_field_transforms(cuda_device, n, dims, r2c, single):
    """three fields through one set of exchanges = three single-field executes, bitwise (same kernels, same order per field);
    ranks as host threads on cuda:0; the 3-D gradient / divergence batch their components on such a plan"""
    torch = torch_mod()
    nprocs = dims[0] * dims[1]
    rdt, cdt = dtypes(single)
    tdt = torch.float32 if single else torch.float64
    es = 4 if single else 8
    nf = 3
    xs = [orc.synthetic_real(n, 70 + f) if r2c else orc.synthetic_complex(n, 70 + f) for f in range(nf)]
    Xs = [orc.forward_r2c(x) if r2c else orc.forward_c2c(x) for x in xs]
    comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
    plans = [acc.Plan(n, comms[r], r2c=r2c, single=single, wire=False, flags=acc.ACCFFT_ESTIMATE, howmany=nf) for r in range(nprocs)]
    if nprocs > 1:
        acc.attach_local_peers(plans)
    assert acc.lib().accfftb_plan_howmany(plans[0]._h) == nf
    streams = [torch.cuda.Stream() for _ in range(nprocs)]
    ins = [[None] * nf for _ in range(nprocs)]
    outs = [[None] * nf for _ in range(nprocs)]
    outs1 = [[None] * nf for _ in range(nprocs)]
    backs = [[None] * nf for _ in range(nprocs)]
    backs1 = [[None] * nf for _ in range(nprocs)]
    grads = [[None] * 3 for _ in range(nprocs)]
    divs = [None] * nprocs
    for r in range(nprocs):
        for f in range(nf):
            loc = orc.local_input(xs[f], n, dims, comms[r].coords).astype(rdt if r2c else cdt)
            t = to_dev(loc, cuda_device) if loc.size else torch.zeros(2, dtype=tdt, device=cuda_device)
            ins[r][f] = torch.view_as_real(t).contiguous() if (not r2c and loc.size) else t
            outs[r][f] = torch.zeros(max(plans[r].alloc_max // es, 2), dtype=tdt, device=cuda_device)
            outs1[r][f] = torch.zeros_like(outs[r][f])
            backs[r][f] = torch.zeros_like(ins[r][f])
            backs1[r][f] = torch.zeros_like(ins[r][f])
        grads[r] = [torch.zeros_like(ins[r][0]) for _ in range(3)]
        divs[r] = torch.zeros_like(ins[r][0])
        plans[r].set_stream(streams[r].cuda_stream)
    torch.cuda.synchronize()

    def body(r):
        p = plans[r]
        with torch.cuda.stream(streams[r]):
            for _ in range(2):
                if r2c:
                    p.execute_r2c_many(ins[r], outs[r])
                    p.execute_c2r_many(outs[r], backs[r])
                    for f in range(nf):
                        p.execute_r2c(ins[r][f], outs1[r][f])
                        p.execute_c2r(outs1[r][f], backs1[r][f])
                else:
                    p.execute_c2c_many(acc.ACCFFT_FORWARD, ins[r], outs[r])
                    p.execute_c2c_many(acc.ACCFFT_BACKWARD, outs[r], backs[r])
                    for f in range(nf):
                        p.execute_c2c(acc.ACCFFT_FORWARD, ins[r][f], outs1[r][f])
                        p.execute_c2c(acc.ACCFFT_BACKWARD, outs1[r][f], backs1[r][f])
            if r2c:
                p.grad_slow(grads[r][0], grads[r][1], grads[r][2], ins[r][0])
                p.divergence_slow(divs[r], ins[r][0], ins[r][1], ins[r][2])

    run_ranks(nprocs, body)
    torch.cuda.synchronize()
    for f in range(nf):
        blocks = []
        for r in range(nprocs):
            osz = plans[r].osize
            blocks.append(outs[r][f].cpu().numpy().view(cdt)[:osz[0] * osz[1] * osz[2]].reshape(osz))
            assert torch.equal(outs[r][f], outs1[r][f]) and torch.equal(backs[r][f], backs1[r][f])
        assert orc.rel_l2(orc.assemble_output(blocks, n, dims, r2c=r2c), Xs[f]) <= TOL[single]
    if r2c:
        tol = 1e-11 if not single else 2e-4
        for ax in range(3):
            bl = [grads[r][ax].cpu().numpy().ravel()[:int(np.prod(plans[r].isize))].reshape(plans[r].isize) for r in range(nprocs)]
            assert orc.rel_l2(orc.assemble_input(bl, n, dims), orc.grad_slow(xs[0], ax)) <= tol
        bl = [divs[r].cpu().numpy().ravel()[:int(np.prod(plans[r].isize))].reshape(plans[r].isize) for r in range(nprocs)]
        assert orc.rel_l2(orc.assemble_input(bl, n, dims), orc.divergence_slow(*xs)) <= tol
    # a plan made for fewer fields refuses the batch
    with pytest.raises(acc.AccfftError):
        q = make_plan((8, 8, 8), True, False)
        try:
            a = torch.zeros(512, dtype=torch.float64, device=cuda_device)
            o = torch.zeros(q.alloc_max // 8, dtype=torch.float64, device=cuda_device)
            q.execute_r2c_many([a, a], [o, o])
        finally:
            q.destroy()
    for p in plans:
        acc.lib().accfftb_plan_destroy(p._h)
        p._h = None
