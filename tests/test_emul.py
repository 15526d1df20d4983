"""Host-logic tests of the plan / stage-descriptor code and the kernel BODIES under the CPU fiber emulator
(tests/emul: test infrastructure; the product never links it).  The emulator executes the same
`stage.cuh` templates the CUDA kernels instantiate, with every rank of the process grid as an OS thread
and peers' buffers addressed directly -- so uneven / empty pencils, the fused pack-on-store addressing and
the barrier placement are validated here without a GPU.  Compared with the numpy oracle; tolerance
1e-12 (double) / 1e-5 (single) relative L2 as BASELINE.json states."""
import os
import sys

import numpy as np
import pytest

from oracle import accfft_oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul_driver as emu  # noqa: E402

CASES = [((8, 8, 8), (1, 1)), ((8, 8, 8), (2, 2)), ((12, 10, 16), (2, 4)), ((6, 12, 10), (3, 2)), ((5, 6, 6), (4, 2)),
         ((3, 5, 4), (4, 4)), ((32, 24, 20), (2, 2)), ((18, 30, 50), (1, 2)), ((45, 15, 90), (2, 1)), ((9, 10, 2), (1, 1)),
         ((75, 4, 36), (1, 1)), ((27, 125, 54), (1, 1)), ((250, 6, 108), (2, 1)), ((4, 135, 250), (1, 2))]


@pytest.fixture(scope="module", autouse=True)
def _build():
    emu.build()


@pytest.mark.parametrize("n,np_", CASES)
def test_r2c_c2r_all_ranks(n, np_):
    x = orc.synthetic_real(n, 7)
    X = orc.forward_r2c(x)
    w = emu.World(n, np_, r2c=True)
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < 1e-12
    assert orc.rel_l2(orc.assemble_output(w.forward(x, padded=True), n, np_), X) < 1e-12     # in-place layout
    assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.spectral(x, "laplace"), n, np_), orc.laplace(x)) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_", CASES[:5])
def test_c2c_all_ranks(n, np_):
    z = orc.synthetic_complex(n, 8)
    Z = orc.forward_c2c(z)
    w = emu.World(n, np_, r2c=False)
    assert orc.rel_l2(orc.assemble_output(w.forward(z), n, np_, r2c=False), Z) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(Z), n, np_), z * z.size) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_", [((8, 8, 8), (1, 1)), ((12, 10, 16), (2, 2)), ((5, 6, 6), (2, 1)), ((9, 5, 10), (1, 2)),
                                   ((6, 12, 10), (3, 2))])
def test_grad_divergence_all_ranks(n, np_):
    # odd lengths included: the full-3-D ("slow") scheme zeroes only X == N/2 (src/operators_cuda.cu:34-37), pinned by
    # tests/golden/ref_*.npz; the per-axis scheme (test_grad_per_axis_*) drops the +-floor(N/2) pair
    x = orc.synthetic_real(n, 17)
    w = emu.World(n, np_, r2c=True)
    g = emu.grad(w, x)
    for ax in range(3):
        assert orc.rel_l2(orc.assemble_input(g[ax], n, np_), orc.grad_slow(x, ax)) < 1e-12, ax
    a = [orc.synthetic_real(n, 20 + i) for i in range(3)]
    dv = emu.divergence(w, *a)
    assert orc.rel_l2(orc.assemble_input(dv, n, np_), orc.divergence_slow(*a)) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.spectral(x, "biharmonic"), n, np_), orc.biharmonic(x)) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_", [((8, 8, 8), (1, 1)), ((12, 10, 16), (2, 2)), ((6, 12, 10), (3, 2))])
@pytest.mark.parametrize("xyz", [0b100, 0b110, 0b101, 0b011, 0b010, 0b001, 0b000])
def test_partial_axis_transforms(n, np_, xyz):
    """std::bitset<3> xyz of accfft_execute_*: unselected axes are only re-laid out (src/accfft_gpu.cpp:477-539)"""
    z = orc.synthetic_complex(n, 11)
    w = emu.World(n, np_, r2c=False)
    Z = orc.forward_partial(z, xyz, r2c=False)
    assert orc.rel_l2(orc.assemble_output(w.forward(z, xyz=xyz), n, np_, r2c=False), Z) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(Z, xyz=xyz), n, np_), orc.inverse_partial(Z, n, xyz, False)) < 1e-12
    w.close()
    if xyz & 4:
        x = orc.synthetic_real(n, 12)
        w = emu.World(n, np_, r2c=True)
        X = orc.forward_partial(x, xyz, r2c=True)
        assert orc.rel_l2(orc.assemble_output(w.forward(x, xyz=xyz), n, np_), X) < 1e-12
        assert orc.rel_l2(orc.assemble_input(w.inverse(X, xyz=xyz), n, np_), orc.inverse_partial(X, n, xyz, True)) < 1e-12
        w.close()


@pytest.mark.parametrize("n,np_", [((12, 10, 16), (2, 2)), ((12, 10, 20), (2, 4)), ((8, 32, 32), (1, 4))])
def test_single_precision(n, np_):
    # (2,4) / (1,4): odd spectral block sizes -> the staged Z-stage stores take their unaligned (element-wise) path
    x = orc.synthetic_real(n, 9)
    w = emu.World(n, np_, r2c=True, single=True)
    X = orc.forward_r2c(x)
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < 1e-5
    assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-5
    w.close()


AXIS_CASES = [((8, 8, 8), (1, 1)), ((6, 10, 8), (1, 1)), ((8, 8, 8), (1, 2)), ((8, 8, 8), (2, 1)), ((12, 10, 16), (2, 4)),
              ((6, 12, 10), (3, 2)), ((5, 6, 6), (4, 2)), ((3, 5, 4), (4, 4)), ((9, 15, 10), (2, 2)), ((32, 24, 20), (2, 2))]


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("n,np_", AXIS_CASES)
def test_single_axis_real_transforms(n, np_, single):
    """accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t (src/accfft_gpu.cpp:1212-1604): per-rank half spectra in the reference's
    osize_xi / osize_yi / osize_0 layouts, round trip = N_axis * x, and the fused per-axis derivative (with and
    without accumulation) -- odd lengths, odd column counts (scalar fallback of the pair accesses), empty ranks."""
    tol = 1e-5 if single else 1e-12
    x = orc.synthetic_real(n, 11)
    y = orc.synthetic_real(n, 12)
    w = emu.World(n, np_, r2c=True, single=single)
    for axis in (0, 1, 2):
        res = emu.axis_r2c(w, x, axis)
        num = sum(float(np.sum(np.abs(a - b) ** 2)) for a, b in res)
        den = sum(float(np.sum(np.abs(b) ** 2)) for a, b in res)
        assert np.sqrt(num / den) < tol, axis
        assert all(a.shape == b.shape for a, b in res)
        assert orc.rel_l2(orc.assemble_input(emu.axis_c2r(w, x, axis), n, np_), x * n[axis]) < tol, axis
        assert orc.rel_l2(orc.assemble_input(emu.axis_deriv(w, x, axis), n, np_), orc.grad(x, axis)) < tol, axis
        acc = orc.assemble_input(emu.axis_deriv(w, x, axis, accumulate_onto=y), n, np_)
        assert orc.rel_l2(acc, orc.grad(x, axis) + y) < tol, axis
    w.close()


@pytest.mark.parametrize("n,np_", [((8, 8, 8), (1, 1)), ((12, 10, 16), (2, 2)), ((5, 6, 6), (2, 1)), ((9, 5, 10), (1, 2)),
                                   ((6, 12, 10), (3, 2)), ((12, 10, 16), (2, 4))])
def test_grad_divergence_per_axis_scheme(n, np_):
    """accfft_grad_gpu_t / accfft_divergence_gpu_t (src/operators_gpu.txx:172-278, :448-553): same numbers as the
    full-3-D ("slow") scheme and the oracle."""
    x = orc.synthetic_real(n, 17)
    w = emu.World(n, np_, r2c=True)
    g = emu.grad(w, x, fast=True)
    for ax in range(3):
        assert orc.rel_l2(orc.assemble_input(g[ax], n, np_), orc.grad(x, ax)) < 1e-12, ax
    gy = emu.grad(w, x, comps=0b010, fast=True)
    assert orc.rel_l2(orc.assemble_input(gy[1], n, np_), orc.grad(x, 1)) < 1e-12
    a = [orc.synthetic_real(n, 20 + i) for i in range(3)]
    dv = emu.divergence(w, *a, fast=True)
    assert orc.rel_l2(orc.assemble_input(dv, n, np_), orc.divergence(*a)) < 1e-12
    # the 3-D path still works afterwards on the same buffers (exchange-buffer discipline)
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), orc.forward_r2c(x)) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_", [((8, 8, 8), (1, 1)), ((12, 10, 16), (2, 2)), ((6, 12, 10), (3, 2)), ((9, 5, 10), (1, 2))])
def test_solver_symbols(n, np_):
    """inv_laplace / inv_biharmonic (CPU-only in the reference, src/operators.txx:275-392) and the general
    (alpha + beta |k|^2)^(+-1) symbol: Poisson filter of steps/step4, heat steps of steps/step6"""
    x = orc.synthetic_real(n, 23)
    w = emu.World(n, np_, r2c=True)
    asm = lambda blocks: orc.assemble_input(blocks, n, np_)
    assert orc.rel_l2(asm(emu.symbol(w, x, "inv_laplace")), orc.inv_laplace(x)) < 1e-12
    assert orc.rel_l2(asm(emu.symbol(w, x, "inv_biharmonic")), orc.inv_biharmonic(x)) < 1e-12
    dt = 0.01
    assert orc.rel_l2(asm(emu.symbol(w, x, "lin", 1.0, -dt)), x + dt * orc.laplace(x)) < 1e-12        # explicit heat step
    assert orc.rel_l2(asm(emu.symbol(w, x, "inv", 1.0, dt)), orc.symbol_solve(x, 1.0, dt, True)) < 1e-12   # implicit heat step
    pois = asm(emu.symbol(w, x, "inv", 0.0, -4 * np.pi ** 2))
    assert orc.rel_l2(pois, orc.symbol_solve(x, 0.0, -4 * np.pi ** 2, True)) < 1e-12                  # Poisson, L = 1
    # solving and re-applying the operator returns the field minus its mean and Nyquist planes
    back = asm(emu.symbol(w, orc.inv_laplace(x), "lin", 0.0, -1.0))
    assert orc.rel_l2(back, orc.laplace(orc.inv_laplace(x))) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_,r2c", [((8, 8, 8), (2, 1), True), ((12, 10, 16), (3, 1), True), ((5, 6, 6), (4, 1), True),
                                       ((16, 12, 6), (2, 1), False), ((32, 24, 20), (4, 1), True), ((12, 10, 16), (2, 2), True)])
@pytest.mark.parametrize("chunks", [2, 3, 7])
def test_pipelined_slab_forward(n, np_, r2c, chunks, monkeypatch):
    """slab grids (P1 == 1): the Z and Y stages run as chunks of i0 (Plan::forward_zy_pipelined); same spectrum.
    (2, 2) is not a slab: the chunk count must be ignored there."""
    monkeypatch.setenv("EMU_PIPELINE_CHUNKS", str(chunks))
    x = orc.synthetic_real(n, 29) if r2c else orc.synthetic_complex(n, 29)
    X = orc.forward_r2c(x) if r2c else orc.forward_c2c(x)
    w = emu.World(n, np_, r2c=r2c)
    for _ in range(2):
        assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_, r2c=r2c), X) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-12
    if r2c:
        assert orc.rel_l2(orc.assemble_output(w.forward(x, padded=True), n, np_), X) < 1e-12
    w.close()


def test_seeded_random_sweep_of_grids_and_process_grids():
    """60 random small (grid, process grid) pairs through the emulated kernels and the real descriptor code: forward
    spectrum, round trip, laplace, per-axis gradient and the single-axis transforms against the oracle.  Lengths are drawn
    from the supported table, process grids up to 4 x 4 (so many ranks own short or empty pencils)."""
    import random
    rng = random.Random(1234)
    small = [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 18, 20, 24, 25, 27, 30]
    for case in range(60):
        n = (rng.choice(small), rng.choice(small), 2 * rng.choice([1, 2, 3, 4, 5, 6, 8, 9, 10, 12]))
        np_ = (rng.choice([1, 2, 3, 4]), rng.choice([1, 2, 3, 4]))
        x = orc.synthetic_real(n, 100 + case)
        w = emu.World(n, np_, r2c=True)
        X = orc.forward_r2c(x)
        assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < 1e-12, (n, np_)
        assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-12, (n, np_)
        if case % 3 == 0:
            assert orc.rel_l2(orc.assemble_input(w.spectral(x, "laplace"), n, np_), orc.laplace(x)) < 1e-12, (n, np_)
            g = emu.grad(w, x, fast=True)
            for ax in range(3):
                ref = orc.grad(x, ax)
                got = orc.assemble_input(g[ax], n, np_)
                den = max(float(np.linalg.norm(ref)), 1e-300)
                assert float(np.linalg.norm(got - ref)) / den < 1e-12 or float(np.linalg.norm(got - ref)) < 1e-12, (n, np_, ax)
            for axis in (0, 1, 2):
                res = emu.axis_r2c(w, x, axis)
                num = sum(float(np.sum(np.abs(a - b) ** 2)) for a, b in res)
                den = sum(float(np.sum(np.abs(b) ** 2)) for a, b in res)
                assert np.sqrt(num / den) < 1e-12, (n, np_, axis)
        w.close()
        if case % 4 == 0:
            z = orc.synthetic_complex(n, 200 + case)
            w = emu.World(n, np_, r2c=False)
            assert orc.rel_l2(orc.assemble_output(w.forward(z), n, np_, r2c=False), orc.forward_c2c(z)) < 1e-12, (n, np_)
            w.close()


def test_seeded_random_sweep_single_precision():
    """the same sweep in single precision (tolerance 1e-5): 25 random grids / process grids, transforms, in-place layout,
    per-axis divergence"""
    import random
    rng = random.Random(4321)
    small = [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 20, 24, 27, 30, 32]
    for case in range(25):
        n = (rng.choice(small), rng.choice(small), 2 * rng.choice([1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 16]))
        np_ = (rng.choice([1, 2, 3, 4]), rng.choice([1, 2, 3, 4]))
        x = orc.synthetic_real(n, 300 + case)
        w = emu.World(n, np_, r2c=True, single=True)
        X = orc.forward_r2c(x)
        assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < 1e-5, (n, np_)
        assert orc.rel_l2(orc.assemble_output(w.forward(x, padded=True), n, np_), X) < 1e-5, (n, np_)
        assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-5, (n, np_)
        if case % 3 == 0:
            a = [orc.synthetic_real(n, 400 + case + i) for i in range(3)]
            ref = orc.divergence(*a)
            got = orc.assemble_input(emu.divergence(w, *a, fast=True), n, np_)
            assert float(np.linalg.norm(got - ref)) <= 2e-5 * max(float(np.linalg.norm(ref)), 1.0), (n, np_)
        w.close()


def test_unsupported_length_is_refused():
    with pytest.raises(RuntimeError):
        emu.World((17, 8, 8), (1, 1), r2c=True)


@pytest.mark.parametrize("chunks,lanes", [(1, 1), (3, 2), (16, 7)])
@pytest.mark.parametrize("n,np_", [((16, 12, 20), (2, 1)), ((12, 10, 16), (4, 1)), ((5, 6, 6), (4, 1)), ((3, 5, 4), (4, 1)),
                                   ((32, 24, 20), (8, 1))])
def test_copy_engine_slab_exchange(monkeypatch, n, np_, chunks, lanes):
    """Plan::forward_slab_ce / inverse_slab_ce (staging buffer + chunked copies instead of peer stores) on slab grids:
    same spectra as the direct path, also when mixed with operations that still use peer stores (operators, single-axis
    transforms) -- the late-A-reader fence of plan.cpp"""
    monkeypatch.setenv("EMU_CE_CHUNKS", str(chunks))
    monkeypatch.setenv("EMU_CE_LANES", str(lanes))
    x = orc.synthetic_real(n, 7)
    X = orc.forward_r2c(x)
    w = emu.World(n, np_, r2c=True)
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < 1e-12
    for _ in range(2):   # inverse after inverse: the second one's copies must wait for the first one's late reads of A
        assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.spectral(x, "laplace"), n, np_), orc.laplace(x)) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-12
    g = emu.grad(w, x, fast=True)
    for ax in range(3):
        assert orc.rel_l2(orc.assemble_input(g[ax], n, np_), orc.grad(x, ax)) < 1e-12
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < 1e-12
    w.close()
    z = orc.synthetic_complex(n, 8)
    Z = orc.forward_c2c(z)
    w = emu.World(n, np_, r2c=False)
    assert orc.rel_l2(orc.assemble_output(w.forward(z), n, np_, r2c=False), Z) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(Z), n, np_), z * z.size) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_", [((7, 14, 22), (1, 1)), ((49, 13, 26), (1, 1)), ((28, 11, 14), (2, 2)), ((13, 7, 98), (1, 2)),
                                   ((56, 44, 28), (2, 1))])
def test_radix_7_11_13_lengths(n, np_):
    """prime radices 7 / 11 / 13 (fft_engine.cuh: DftPrime) and the composite 14"""
    x = orc.synthetic_real(n, 71)
    X = orc.forward_r2c(x)
    w = emu.World(n, np_, r2c=True)
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < 1e-12
    w.close()
    z = orc.synthetic_complex(n, 72)
    Z = orc.forward_c2c(z)
    w = emu.World(n, np_, r2c=False)
    assert orc.rel_l2(orc.assemble_output(w.forward(z), n, np_, r2c=False), Z) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(Z), n, np_), z * z.size) < 1e-12
    w.close()


@pytest.mark.parametrize("planes", [1, 3, 5])
@pytest.mark.parametrize("n,r2c", [((16, 12, 20), True), ((7, 10, 16), True), ((12, 8, 6), False)])
def test_l2_chunked_local_passes(monkeypatch, n, r2c, planes):
    """Plan::forward_zy_l2 / inverse_yz_l2: Z and Y chunk by chunk over i0 through the reused head of A / B (single rank)"""
    monkeypatch.setenv("EMU_L2_PLANES", str(planes))
    x = orc.synthetic_real(n, 5) if r2c else orc.synthetic_complex(n, 5)
    X = orc.forward_r2c(x) if r2c else orc.forward_c2c(x)
    w = emu.World(n, (1, 1), r2c=r2c)
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, (1, 1), r2c=r2c), X) < 1e-12
    assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, (1, 1)), x * x.size) < 1e-12
    if r2c:
        assert orc.rel_l2(orc.assemble_output(w.forward(x, padded=True), n, (1, 1)), X) < 1e-12
        assert orc.rel_l2(orc.assemble_input(w.inverse(X, padded=True), n, (1, 1)), x * x.size) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_,r2c", [((12, 10, 16), (2, 2), True), ((5, 6, 6), (4, 1), True), ((3, 5, 4), (4, 4), True), ((8, 8, 8), (1, 1), True),
                                       ((16, 12, 20), (2, 4), True), ((6, 10, 5), (2, 2), False), ((9, 8, 7), (3, 1), False)])
def test_batched_fields_share_the_exchanges(monkeypatch, n, np_, r2c):
    """accfftb_execute_*_many (Plan::execute_forward_many / inverse_many): three fields through one set of barriers, each in its
    own field set of the exchange buffers -- results equal three single-field transforms; and the full-3-D gradient /
    divergence, which batch their three components on such a plan, still match the oracle"""
    monkeypatch.setenv("EMU_BATCH", "3")
    w = emu.World(n, np_, r2c=r2c)
    xs = [orc.synthetic_real(n, 60 + i) if r2c else orc.synthetic_complex(n, 60 + i) for i in range(3)]
    Xs = [orc.forward_r2c(x) if r2c else orc.forward_c2c(x) for x in xs]
    for rep in range(2):   # twice: buffer-reuse ordering across batched operations
        res = w.forward_many(xs)
        for f in range(3):
            assert orc.rel_l2(orc.assemble_output(res[f], n, np_, r2c=r2c), Xs[f]) < 1e-12, f
        back = w.inverse_many(Xs)
        for f in range(3):
            assert orc.rel_l2(orc.assemble_input(back[f], n, np_), xs[f] * xs[f].size) < 1e-12, f
    two = w.forward_many(xs[:2])   # fewer fields than the plan holds
    assert orc.rel_l2(orc.assemble_output(two[1], n, np_, r2c=r2c), Xs[1]) < 1e-12
    # single-field calls on a batched plan use field set 0
    assert orc.rel_l2(orc.assemble_output(w.forward(xs[2]), n, np_, r2c=r2c), Xs[2]) < 1e-12
    if r2c:
        g = emu.grad(w, xs[0])
        for ax in range(3):
            assert orc.rel_l2(orc.assemble_input(g[ax], n, np_), orc.grad_slow(xs[0], ax)) < 1e-12, ax
        dv = emu.divergence(w, *xs)
        assert orc.rel_l2(orc.assemble_input(dv, n, np_), orc.divergence_slow(*xs)) < 1e-12
    w.close()


@pytest.mark.parametrize("n,np_,single", [((6, 10, 384), (1, 1), False), ((12, 8, 384), (2, 2), False), ((6, 10, 384), (1, 2), True),
                                          ((8, 6, 512), (1, 1), True), ((10, 12, 256), (2, 1), True), ((4, 6, 1024), (1, 1), True),
                                          ((6, 8, 768), (1, 1), True)])
def test_contiguous_kernels_run_the_gpu_builds_own_schedules(n, np_, single):
    """schedules.def gives the contiguous (Z pass) kernels their own schedules for some lengths (KEEP_C: a warp per pencil at 192
    points; KEEP_FC: 16 values per thread in single precision); tests/emul/emul_alt.cpp instantiates exactly those lines, so the
    Z stages below run what the GPU runs -- incl. passes in which part of a pencil's threads idle"""
    assert emu.lib().emu_alt_schedule_count() >= 9
    before = emu.lib().emu_alt_launches()
    x = orc.synthetic_real(n, 33)
    X = orc.forward_r2c(x)
    w = emu.World(n, np_, r2c=True, single=single)
    tol = 1e-5 if single else 1e-12
    assert orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X) < tol
    assert orc.rel_l2(orc.assemble_input(w.inverse(X), n, np_), x * x.size) < tol
    w.close()
    assert emu.lib().emu_alt_launches() > before, "the Z stages did not pick a kind-specific schedule"
    z = orc.synthetic_complex((n[0], n[1], n[2] // 2), 34)
    w = emu.World((n[0], n[1], n[2] // 2), np_, r2c=False, single=single)
    assert orc.rel_l2(orc.assemble_output(w.forward(z), z.shape, np_, r2c=False), orc.forward_c2c(z)) < tol
    w.close()
