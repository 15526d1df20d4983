"""C-ABI boundary checks that need no GPU: the library loads, exports every symbol include/accfft_b200.h
declares, and its host-side integer math (decomposition, communicator dims) equals the oracle's restatement
of src/accfft_common.cpp.  No compute entry point is called here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import accfft_b200 as acc
from oracle import accfft_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "accfft_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(accfftb_[a-z0-9_]+)\s*\(", src)))


def test_library_present_and_loads():
    assert os.path.exists(acc.lib_path()), "build with `make -C accfft_b200/csrc -j8`"
    assert acc.lib().accfftb_version().startswith(b"accfft_b200")


def test_every_declared_symbol_is_exported():
    syms = declared_symbols()
    assert len(syms) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", acc.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    L = acc.lib()
    for s in syms:
        assert getattr(L, s) is not None


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", acc.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import accfft_b200.api as api
    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setenv("ACCFFT_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(acc.AccfftError):
        api.lib()
    monkeypatch.delenv("ACCFFT_B200_LIB")
    monkeypatch.setattr(api, "_LIB", None)
    api.lib()


GRIDS = [((512, 512, 512), (1, 1)), ((512, 512, 512), (2, 4)), ((512, 512, 512), (2, 1)), ((512, 512, 512), (2, 2)),
         ((1024, 1024, 1024), (2, 4)), ((768, 768, 384), (2, 4)), ((128, 128, 128), (1, 2)), ((12, 10, 14), (2, 4)),
         ((6, 9, 10), (3, 2)), ((5, 7, 6), (4, 2)), ((3, 5, 4), (4, 4)), ((16, 12, 6), (2, 4)), ((7, 3, 2), (8, 2))]


@pytest.mark.parametrize("n,dims", GRIDS)
@pytest.mark.parametrize("sizeof_real", [8, 4])
def test_local_sizes_match_oracle(n, dims, sizeof_real):
    nprocs = dims[0] * dims[1]
    for rank in range(nprocs):
        comm = acc.Comm(rank, nprocs, dims)
        assert comm.coords == orc.rank_to_coords(rank, dims)
        fn = acc.accfft_local_size_dft_r2c_gpu if sizeof_real == 8 else acc.accfft_local_size_dft_r2c_gpuf
        alloc, isize, istart, osize, ostart = fn(n, comm)
        want = orc.local_size_dft_r2c(n, dims, comm.coords, sizeof_real)
        assert (alloc, isize, istart, osize, ostart) == (want["alloc_max"], want["isize"], want["istart"],
                                                         want["osize"], want["ostart"])
        fn = acc.accfft_local_size_dft_c2c_gpu if sizeof_real == 8 else acc.accfft_local_size_dft_c2c_gpuf
        alloc, isize, istart, osize, ostart = fn(n, comm)
        want = orc.local_size_dft_c2c(n, dims, comm.coords, sizeof_real)
        assert (alloc, isize, istart, osize, ostart) == (want["alloc_max"], want["isize"], want["istart"],
                                                         want["osize"], want["ostart"])


@pytest.mark.parametrize("n,dims", GRIDS[:8])
def test_stage_sizes_match_oracle(n, dims):
    L = acc.lib()
    i3 = lambda v: (ctypes.c_int * 3)(*v)
    i2 = lambda v: (ctypes.c_int * 2)(*v)
    names = ["osize_0", "osize_1", "osize_2", "osize_y", "osize_yi", "osize_x", "osize_xi"]
    starts = ["ostart_0", "ostart_1", "ostart_2", "ostart_y", None, None, None]
    for rank in range(dims[0] * dims[1]):
        coords = orc.rank_to_coords(rank, dims)
        want = orc.local_size_dft_r2c(n, dims, coords, 8)
        for which, name in enumerate(names):
            size, start = i3((0,) * 3), i3((0,) * 3)
            assert L.accfftb_stage_size(i3(n), i2(dims), i2(coords), 0, 8, which, size, start) == 0
            assert list(size) == list(want[name]), (name, rank)
            if starts[which]:
                assert list(start) == list(want[starts[which]]), (starts[which], rank)


def test_sizes_match_oracle_on_a_seeded_random_sweep():
    """300 random (grid, process grid) pairs, every rank: r2c / c2c size queries and all seven stage layouts of the C ABI
    (64-bit decomposition in csrc/decomp.hpp) against the numpy restatement of src/accfft_common.cpp:49-256 -- uneven,
    short and empty blocks included (more ranks than planes along an axis)."""
    import random
    rng = random.Random(20240607)
    L = acc.lib()
    i3 = lambda v: (ctypes.c_int * 3)(*v)
    i2 = lambda v: (ctypes.c_int * 2)(*v)
    names = ["osize_0", "osize_1", "osize_2", "osize_y", "osize_yi", "osize_x", "osize_xi"]
    for _ in range(300):
        n = tuple(rng.choice([1, 2, 3, 5, 6, 8, 12, 17, 30, 64, 100, 257]) for _ in range(3))
        dims = (rng.choice([1, 2, 3, 4, 8]), rng.choice([1, 2, 3, 4, 8]))
        nprocs = dims[0] * dims[1]
        for rank in rng.sample(range(nprocs), min(nprocs, 4)):
            comm = acc.Comm(rank, nprocs, dims)
            for sizeof_real, r2c_fn, c2c_fn in ((8, acc.accfft_local_size_dft_r2c_gpu, acc.accfft_local_size_dft_c2c_gpu),
                                                (4, acc.accfft_local_size_dft_r2c_gpuf, acc.accfft_local_size_dft_c2c_gpuf)):
                got = r2c_fn(n, comm)
                want = orc.local_size_dft_r2c(n, dims, comm.coords, sizeof_real)
                assert got == (want["alloc_max"], want["isize"], want["istart"], want["osize"], want["ostart"]), (n, dims, rank)
                got = c2c_fn(n, comm)
                wc = orc.local_size_dft_c2c(n, dims, comm.coords, sizeof_real)
                assert got == (wc["alloc_max"], wc["isize"], wc["istart"], wc["osize"], wc["ostart"]), (n, dims, rank)
            want = orc.local_size_dft_r2c(n, dims, comm.coords, 8)
            for which, name in enumerate(names):
                size, start = i3((0,) * 3), i3((0,) * 3)
                assert L.accfftb_stage_size(i3(n), i2(dims), i2(comm.coords), 0, 8, which, size, start) == 0
                assert list(size) == list(want[name]), (n, dims, rank, name)


def test_overflow_guard_returns_minus_one():
    # src/accfft_common.cpp:179-184: isize product > 2^31-1 -> -1
    comm = acc.Comm(0, 1, (1, 1))
    alloc, *_ = acc.accfft_local_size_dft_r2c_gpu((2048, 2048, 1024), comm)
    assert alloc == -1
    alloc, *_ = acc.accfft_local_size_dft_r2c_gpu((1024, 1024, 1024), comm)
    assert alloc == 2 * 1024 * 1024 * 513 * 8


@pytest.mark.parametrize("nprocs,dims_in,dims_out", [(8, [0, 0], [4, 2]), (8, [2, 4], [2, 4]), (2, [0, 0], [2, 1]),
                                                     (4, [1, 1], [2, 2]), (6, [0, 0], [3, 2]), (1, [0, 0], [1, 1]),
                                                     (8, [8, 1], [8, 1]), (7, [2, 2], [7, 1])])
def test_create_comm_overwrites_bad_dims(nprocs, dims_in, dims_out):
    # accfft_create_comm silently overwrites c_dims (src/accfft_common.cpp:281-289)
    d = list(dims_in)
    comm = acc.accfft_create_comm(d, rank=0, nprocs=nprocs)
    assert d == dims_out and list(comm.c_dims) == dims_out
    assert tuple(dims_out) == orc.dims_create(nprocs) or dims_in == dims_out


def test_supported_lengths():
    for n in (2, 3, 4, 5, 6, 8, 12, 16, 27, 64, 96, 125, 128, 192, 216, 250, 256, 384, 500, 512, 768, 1000, 1024, 1728, 2000, 2048,
              7, 11, 13, 14, 49, 98, 448, 896, 1792, 44, 52, 4096):
        assert acc.length_fast(n) and acc.length_supported(n), n
    for n in (81, 625, 1023, 17, 77, 343, 7168):     # no compiled schedule: planned through the generic O(n^2) stage kernel
        assert not acc.length_fast(n) and acc.length_supported(n), n
    for n in (0, -4, 7169, 8192):
        assert not acc.length_supported(n), n


def test_host_tensor_is_rejected_not_computed():
    """the product has no CPU path: host tensors are refused before anything is launched"""
    import torch
    from accfft_b200.api import _ptr
    with pytest.raises(acc.AccfftError):
        _ptr(torch.zeros(8))
