#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz from the REFERENCE'S OWN CPU library run here.

oracle/Makefile (target `ref`) compiles /root/reference/src/{accfft,accfftf,accfft_common,transpose,operators,operatorsf}.cpp
unmodified against oracle/shim/fftw3.h (definitional long-double DFT) and include/mpi_shim/mpi.h (single-node MPI subset)
together with oracle/ref_driver.cpp into oracle/_ref/ref_driver.  This script runs that binary under accfft_mpirun for
each case below (one process per rank) and packs what every rank dumped -- decomposition (isize / istart / osize / ostart /
alloc_max), forward spectrum block, c2r round trip, grad / laplace / divergence / biharmonic / inv_laplace /
inv_biharmonic blocks -- into one .npz per case.  Inputs are regenerated from the seed (oracle.hashed_field), so they are
not stored.  Needs /root/reference (development container only); the fixtures are committed.

    python tests/golden/make_ref_golden.py            # rebuild every fixture
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# (name, kind, n, c_dims, seed, flags)
CASES = [
    ("r2c_12x10x14_2x4", "r2c", (12, 10, 14), (2, 4), 11, "estimate"),          # uneven pencils in both exchanges
    ("r2c_12x10x14_2x4_measure", "r2c", (12, 10, 14), (2, 4), 11, "measure"),   # the tuner picks another transpose variant
    ("r2c_5x7x6_4x2", "r2c", (5, 7, 6), (4, 2), 12, "estimate"),                # short last blocks
    ("r2c_3x5x4_4x4", "r2c", (3, 5, 4), (4, 4), 13, "estimate"),                # empty ranks (P0 > N0)
    ("r2c_16x12x20_1x1", "r2c", (16, 12, 20), (1, 1), 14, "estimate"),
    ("r2c_16x12x20_2x1", "r2c", (16, 12, 20), (2, 1), 14, "estimate"),          # slab
    ("r2c_16x12x20_1x2", "r2c", (16, 12, 20), (1, 2), 14, "estimate"),          # the reference's step1 default grid
    ("r2c_16x12x20_4x2", "r2c", (16, 12, 20), (4, 2), 14, "estimate"),
    ("r2c_inplace_12x10x14_2x2", "r2c_inplace", (12, 10, 14), (2, 2), 15, "estimate"),
    ("r2c_6x5x7_2x2", "r2c", (6, 5, 7), (2, 2), 16, "estimate"),                # odd N2
    ("c2c_16x12x20_2x4", "c2c", (16, 12, 20), (2, 4), 17, "estimate"),
    ("c2c_6x10x5_2x2", "c2c", (6, 10, 5), (2, 2), 18, "estimate"),
    ("c2c_9x8x7_3x1", "c2c", (9, 8, 7), (3, 1), 19, "estimate"),
]

R2C_FIELDS = ("forward", "roundtrip", "grad_x", "grad_y", "grad_z", "grad_slow_x", "grad_slow_y", "grad_slow_z", "div_grad", "div_slow_grad_slow",
              "laplace", "biharmonic", "inv_laplace", "inv_biharmonic")


def run_case(name, kind, n, dims, seed, flags, driver, mpirun):
    nprocs = dims[0] * dims[1]
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [mpirun, "-n", str(nprocs), "--gpus-per-node", "0", driver, tmp, name, kind, *map(str, n), *map(str, dims),
               str(seed), flags]
        env = dict(os.environ, OMP_NUM_THREADS="1")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        if r.returncode != 0:
            raise RuntimeError("%s failed (%d):\n%s\n%s" % (name, r.returncode, r.stdout[-2000:], r.stderr[-2000:]))
        out = {"n": np.array(n), "c_dims": np.array(dims), "seed": np.array(seed), "kind": np.array(kind)}
        for rank in range(nprocs):
            pre = os.path.join(tmp, "%s.r%d." % (name, rank))
            meta = np.fromfile(pre + "meta.bin", dtype=np.int32)
            out["r%d_meta" % rank] = meta
            isz, osz = tuple(meta[1:4]), tuple(meta[7:10])
            if kind == "c2c":
                out["r%d_forward" % rank] = np.fromfile(pre + "forward.bin", dtype=np.complex128).reshape(osz)
                out["r%d_backward" % rank] = np.fromfile(pre + "backward.bin", dtype=np.complex128).reshape(isz)
            else:
                for f in R2C_FIELDS:
                    path = pre + f + ".bin"
                    if not os.path.exists(path):
                        continue   # in-place cases dump the transforms only
                    if f == "forward":
                        out["r%d_%s" % (rank, f)] = np.fromfile(path, dtype=np.complex128).reshape(osz)
                    else:
                        out["r%d_%s" % (rank, f)] = np.fromfile(path, dtype=np.float64).reshape(isz)
        return out


def main():
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    mpirun = os.path.join(ROOT, "accfft_b200", "bin", "accfft_mpirun")
    if not os.path.exists(driver):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    for case in CASES:
        data = run_case(*case, driver, mpirun)
        path = os.path.join(HERE, "ref_%s.npz" % case[0])
        np.savez_compressed(path, **data)
        print("wrote %s (%d arrays, %d bytes)" % (os.path.relpath(path, ROOT), len(data), os.path.getsize(path)))


if __name__ == "__main__":
    main()
