"""The CUDA path against outputs of the REFERENCE'S OWN CPU library (tests/golden/ref_*.npz; generator
tests/golden/make_ref_golden.py; CPU twin of this file: tests/test_ref_golden.py).

Every fixture is replayed on the same process grid with one host thread per rank on cuda:0 (peer-addressed
exchange + inter-GPU barrier kernels), through the C ABI; every rank's block -- decomposition, forward spectrum,
c2r round trip, grad / grad_slow / divergence / laplace / biharmonic / inv_laplace / inv_biharmonic -- must match what
that rank of the reference produced, relative L2 <= 1e-12 (BASELINE.json's double-precision tolerance).
"""
import glob
import os
import threading

import numpy as np
import pytest

import accfft_b200 as acc
from oracle import accfft_oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
TOL = 1e-12


def run_ranks(nprocs, fn):
    import torch
    errs = []

    def wrap(r):
        try:
            torch.cuda.set_device(0)
            fn(r)
        except Exception as e:  # pragma: no cover
            errs.append((r, repr(e)))

    th = [threading.Thread(target=wrap, args=(r,)) for r in range(nprocs)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=180)
    assert not errs, errs
    assert not any(t.is_alive() for t in th), "rank thread hung"


def close(got, want, what):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if want.size:
        err = orc.rel_l2(got, want)
        assert err <= TOL, "%s: rel-L2 %.3e vs the reference's own output" % (what, err)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[4:-4] for f in FILES])
def test_cuda_path_matches_reference_run(cuda_device, path):
    import torch
    z = np.load(path)
    n = tuple(int(v) for v in z["n"])
    dims = tuple(int(v) for v in z["c_dims"])
    seed, kind = int(z["seed"]), str(z["kind"])
    r2c = kind != "c2c"
    inplace = kind == "r2c_inplace"
    nprocs = dims[0] * dims[1]
    assert all(acc.length_supported(v) for v in n)   # odd N2 and unscheduled lengths run the generic stage kernel
    x = orc.hashed_field(n, seed, c2c=not r2c)
    comms = [acc.Comm(r, nprocs, dims) for r in range(nprocs)]
    plans = [acc.Plan(n, comms[r], r2c=r2c, single=False, wire=False) for r in range(nprocs)]
    if nprocs > 1:
        acc.attach_local_peers(plans)
    streams = [torch.cuda.Stream() for _ in range(nprocs)]
    res = [dict() for _ in range(nprocs)]
    n2c = n[2] // 2 + 1
    for r in range(nprocs):
        meta = [int(v) for v in z["r%d_meta" % r]]
        p = plans[r]
        # decomposition through the C ABI vs the reference's accfft_local_size_dft_*  (src/accfft_common.cpp:98-256)
        assert meta[0] == p.alloc_max and meta[1:4] == list(p.isize) and meta[4:7] == list(p.istart)
        assert meta[7:10] == list(p.osize) and meta[10:13] == list(p.ostart)
        p.set_stream(streams[r].cuda_stream)

    def body(r):
        p = plans[r]
        isz, osz = tuple(p.isize), tuple(p.osize)
        nin, nout = int(np.prod(isz)), int(np.prod(osz))
        loc = orc.local_input(x, n, dims, comms[r].coords)
        with torch.cuda.stream(streams[r]):
            out = torch.zeros(max(p.alloc_max // 8, 2), dtype=torch.float64, device=cuda_device)
            if not r2c:
                xd = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(loc)).to(cuda_device)).contiguous() if nin else \
                    torch.zeros(2, dtype=torch.float64, device=cuda_device)
                back = torch.zeros_like(xd)
                p.execute_c2c(acc.ACCFFT_FORWARD, xd, out)
                res[r]["forward"] = out.cpu().numpy().view(np.complex128)[:nout].reshape(osz)
                p.execute_c2c(acc.ACCFFT_BACKWARD, out, back)
                res[r]["backward"] = back.cpu().numpy().ravel().view(np.complex128)[:nin].reshape(isz)
                return
            if inplace:   # one alloc_max buffer, rows padded to 2 * (N2/2 + 1) reals (steps/step2/step2_gpu.cpp:41-48)
                buf = np.zeros(max(p.alloc_max // 8, 2))
                if nin:
                    pad = np.zeros(isz[:2] + (2 * n2c,))
                    pad[:, :, :n[2]] = loc
                    buf[:pad.size] = pad.ravel()
                bd = torch.from_numpy(buf).to(cuda_device)
                p.execute_r2c(bd, bd)
                res[r]["forward"] = bd.cpu().numpy().view(np.complex128)[:nout].reshape(osz)
                p.execute_c2r(bd, bd)
                rt = bd.cpu().numpy()[:isz[0] * isz[1] * 2 * n2c].reshape(isz[:2] + (2 * n2c,))[:, :, :n[2]]
                res[r]["roundtrip"] = np.ascontiguousarray(rt)
                return
            xd = torch.from_numpy(np.ascontiguousarray(loc)).to(cuda_device) if nin else \
                torch.zeros(2, dtype=torch.float64, device=cuda_device)
            p.execute_r2c(xd, out)
            res[r]["forward"] = out.cpu().numpy().view(np.complex128)[:nout].reshape(osz)
            back = torch.zeros_like(xd)
            p.execute_c2r(out, back)
            res[r]["roundtrip"] = back.cpu().numpy()[:nin].reshape(isz)

            def grab(t):
                return t.cpu().numpy()[:nin].reshape(isz)

            g = [torch.zeros_like(xd) for _ in range(3)]
            p.grad(g[0], g[1], g[2], xd)
            for ax, nm in enumerate("xyz"):
                res[r]["grad_" + nm] = grab(g[ax])
            w = torch.zeros_like(xd)
            p.divergence(w, g[0], g[1], g[2])
            res[r]["div_grad"] = grab(w)
            for nm, fn in (("laplace", p.laplace), ("biharmonic", p.biharmonic), ("inv_laplace", p.inv_laplace),
                           ("inv_biharmonic", p.inv_biharmonic)):
                fn(w, xd)
                res[r][nm] = grab(w)
            p.grad_slow(g[0], g[1], g[2], xd)
            for ax, nm in enumerate("xyz"):
                res[r]["grad_slow_" + nm] = grab(g[ax])
            p.divergence_slow(w, g[0], g[1], g[2])
            res[r]["div_slow_grad_slow"] = grab(w)

    run_ranks(nprocs, body)
    torch.cuda.synchronize()
    for r in range(nprocs):
        assert res[r], "rank %d produced nothing" % r
        for name, got in res[r].items():
            close(got, z["r%d_%s" % (r, name)], "%s on rank %d of %s" % (name, r, os.path.basename(path)))
    for p in plans:
        acc.lib().accfftb_plan_destroy(p._h)
        p._h = None
