"""The numpy oracle against outputs of the REFERENCE'S OWN CPU library (tests/golden/ref_*.npz).

The fixtures come from /root/reference/src/{accfft,accfft_common,transpose,operators}.cpp compiled unmodified and run here
as 1..16 ranks (tests/golden/make_ref_golden.py, oracle/Makefile target `ref`): decomposition, forward spectra, round
trips and every spectral operator, per rank.  This is the pin of oracle/accfft_oracle.py; the GPU twin of
this file is tests/test_gpu_parity.py::test_ref_golden_*.
"""
import glob
import os

import numpy as np
import pytest

from oracle import accfft_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
TOL = 1e-13   # both sides are double precision evaluations of the same sums


def load(path):
    z = np.load(path)
    return z, tuple(int(v) for v in z["n"]), tuple(int(v) for v in z["c_dims"]), int(z["seed"]), str(z["kind"])


def test_fixtures_present():
    assert len(FILES) >= 12


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[4:-4] for f in FILES])
def test_decomposition_matches_reference(path):
    """isize / istart / osize / ostart / alloc_max of every rank: src/accfft_common.cpp:49-192, :203-256"""
    z, n, dims, seed, kind = load(path)
    for rank in range(dims[0] * dims[1]):
        meta = [int(v) for v in z["r%d_meta" % rank]]
        co = orc.rank_to_coords(rank, dims)
        L = orc.local_size_dft_c2c(n, dims, co) if kind == "c2c" else orc.local_size_dft_r2c(n, dims, co)
        assert meta[0] == L["alloc_max"]
        assert meta[1:4] == list(L["isize"]) and meta[4:7] == list(L["istart"])
        assert meta[7:10] == list(L["osize"]) and meta[10:13] == list(L["ostart"])
        assert meta[13:16] == [dims[0], dims[1], dims[0] * dims[1]]


def _check(z, rank, name, want):
    got = z["r%d_%s" % (rank, name)]
    assert got.shape == want.shape, (name, rank, got.shape, want.shape)
    if got.size:
        err = orc.rel_l2(got, want)
        assert err <= TOL, "%s on rank %d: rel-L2 %.3e" % (name, rank, err)


@pytest.mark.parametrize("path", [f for f in FILES if "ref_c2c" not in f], ids=lambda f: os.path.basename(f)[4:-4])
def test_r2c_transforms_and_operators_match_reference(path):
    z, n, dims, seed, kind = load(path)
    x = orc.hashed_field(n, seed)
    X = orc.forward_r2c(x)
    fields = {"roundtrip": x * x.size}
    if kind == "r2c":
        g = [orc.grad(x, ax) for ax in range(3)]
        gs = [orc.grad_slow(x, ax) for ax in range(3)]
        fields.update({"grad_x": g[0], "grad_y": g[1], "grad_z": g[2], "grad_slow_x": gs[0], "grad_slow_y": gs[1],
                       "grad_slow_z": gs[2], "div_grad": orc.divergence(*g), "div_slow_grad_slow": orc.divergence_slow(*gs),
                       "laplace": orc.laplace(x),
                       "biharmonic": orc.biharmonic(x), "inv_laplace": orc.inv_laplace(x),
                       "inv_biharmonic": orc.inv_biharmonic(x)})
    for rank in range(dims[0] * dims[1]):
        co = orc.rank_to_coords(rank, dims)
        _check(z, rank, "forward", orc.local_output(X, n, dims, co, r2c=True))
        for name, glob_field in fields.items():
            _check(z, rank, name, orc.local_input(glob_field, n, dims, co))


@pytest.mark.parametrize("path", [f for f in FILES if "ref_c2c" in f], ids=lambda f: os.path.basename(f)[4:-4])
def test_c2c_transforms_match_reference(path):
    z, n, dims, seed, kind = load(path)
    x = orc.hashed_field(n, seed, c2c=True)
    X = orc.forward_c2c(x)
    for rank in range(dims[0] * dims[1]):
        co = orc.rank_to_coords(rank, dims)
        _check(z, rank, "forward", orc.local_output(X, n, dims, co, r2c=False))
        _check(z, rank, "backward", orc.local_input(x * x.size, n, dims, co))


def test_measure_and_estimate_agree():
    """the reference's tuned transpose variant (ACCFFT_MEASURE) and its default one move the same data"""
    a = np.load(os.path.join(GOLDEN, "ref_r2c_12x10x14_2x4.npz"))
    b = np.load(os.path.join(GOLDEN, "ref_r2c_12x10x14_2x4_measure.npz"))
    for k in a.files:
        if k.endswith("_forward") or k.endswith("_roundtrip"):
            assert np.array_equal(a[k], b[k]), k


# ---- the same kernel bodies (stage.cuh / plan.cpp under the CPU fiber emulator) against the reference-run fixtures ----
def _emu():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
    import emul_driver
    return emul_driver


def _emu_supported(n, kind):
    import accfft_b200 as acc
    if kind != "c2c" and n[2] % 2:
        return False
    lens = (n[0], n[1], n[2] // 2 if kind != "c2c" else n[2])
    return all(acc.length_fast(v) for v in lens)   # (the emulator has no generic O(N^2) kernel)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[4:-4] for f in FILES])
def test_emulated_kernels_match_reference_run(path):
    """plan.cpp + the stage kernel bodies, all ranks as emulator threads, vs what each rank of the reference produced"""
    z, n, dims, seed, kind = load(path)
    if not _emu_supported(n, kind):
        pytest.skip("length / odd-N2 r2c not supported by the kernel schedules")
    emu = _emu()
    r2c = kind != "c2c"
    x = orc.hashed_field(n, seed, c2c=not r2c)
    w = emu.World(n, dims, r2c=r2c)
    nprocs = dims[0] * dims[1]

    def check(blocks, name):
        for rank in range(nprocs):
            want = z["r%d_%s" % (rank, name)]
            got = np.asarray(blocks[rank]).reshape(want.shape)
            if want.size:
                err = orc.rel_l2(got, want)
                assert err <= 1e-12, "%s on rank %d: rel-L2 %.3e" % (name, rank, err)

    fwd = w.forward(x, padded=(kind == "r2c_inplace"))
    check(fwd, "forward")
    X = orc.assemble_output(fwd, n, dims, r2c=r2c)
    check(w.inverse(X, padded=(kind == "r2c_inplace")), "roundtrip" if r2c else "backward")
    if kind == "r2c":
        check(w.spectral(x, "laplace"), "laplace")
        check(w.spectral(x, "biharmonic"), "biharmonic")
        check(emu.symbol(w, x, "inv_laplace"), "inv_laplace")
        check(emu.symbol(w, x, "inv_biharmonic"), "inv_biharmonic")
        g = emu.grad(w, x, fast=True)
        gs = emu.grad(w, x, fast=False)
        for ax, nm in enumerate("xyz"):
            check(g[ax], "grad_" + nm)
            check(gs[ax], "grad_slow_" + nm)
        gg = [orc.assemble_input(g[ax], n, dims) for ax in range(3)]
        check(emu.divergence(w, *gg, fast=True), "div_grad")
        ggs = [orc.assemble_input(gs[ax], n, dims) for ax in range(3)]
        check(emu.divergence(w, *ggs, fast=False), "div_slow_grad_slow")
    w.close()
