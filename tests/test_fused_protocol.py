"""CPU model of the fused two-pass kernel's scheduler (accfft_b200/csrc/fused.cuh) on the real work-list arithmetic
(csrc/fused_sched.hpp): randomised interleavings of 1 ... 600 simulated CTAs must keep every read-after-write and write-after-read
dependency on the L2 ring, run every tile exactly once and never deadlock.  The kernel itself is checked bitwise against the
two-kernel path on the GPU (tests/test_gpu_parity.py::test_fused_local_pair_*)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul_driver as emu  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _build():
    emu.build()


# (planes, pencils per plane, pencils per Z tile, Y tiles per plane): 512^3 / 256^3 double r2c, a ragged small grid, 8-GPU slab share
SHAPES = [(512, 512, 16, 33), (256, 256, 16, 17), (37, 64, 16, 5), (64, 512, 16, 33), (3, 512, 16, 33), (1, 64, 16, 5)]


@pytest.mark.parametrize("zfirst", [True, False])
@pytest.mark.parametrize("nctas", [1, 2, 7, 296, 592])
@pytest.mark.parametrize("shape", SHAPES)
def test_default_parameters_hold_every_dependency(shape, nctas, zfirst):
    nA, nBz, pcz, tpa = shape
    for seed in (1, 2):
        rc, sh = emu.fused_protocol_sim(nA, nBz, pcz, tpa, nctas, zfirst=zfirst, seed=seed)
        assert rc == 0, (rc, sh)
        assert sh["ns"] > sh["lag"] and sh["ns"] & (sh["ns"] - 1) == 0 and sh["cp"] & (sh["cp"] - 1) == 0


@pytest.mark.parametrize("cp,lag,ns", [(1, 1, 2), (2, 1, 2), (4, 3, 4), (1, 8, 16), (8, 1, 2)])
@pytest.mark.parametrize("flags", [0, 1, 2, 3])
def test_tight_rings_and_debug_modes(cp, lag, ns, flags):
    """two-slot rings with the second pass one chunk behind the first: nearly every tile has to wait; and the kernel's debugging
    modes (1 = signal at once, 2 = no early check)"""
    for nctas in (1, 3, 40, 296):
        for zfirst in (True, False):
            rc, sh = emu.fused_protocol_sim(64, 64, 16, 5, nctas, cp=cp, lag=lag, ns=ns, zfirst=zfirst, seed=7 + nctas, flags=flags)
            assert rc == 0, (rc, sh, nctas, zfirst)
            assert sh["lag"] == lag and sh["cp"] == cp and sh["ring_planes"] == (sh["ns"] * cp if sh["ns"] * cp < 64 else 0)


def test_the_model_catches_a_broken_ring():
    """sanity of the checker itself: without synchronisation (the kernel's timing-only flag 4 has no counterpart here, so the
    model is asked for a ring with fewer slots than the lag) the parameter routine must refuse or repair, never produce ns <= lag"""
    rc, sh = emu.fused_protocol_sim(64, 64, 16, 5, 40, cp=1, lag=5, ns=2)
    assert rc == 0 and sh["ns"] > sh["lag"]
