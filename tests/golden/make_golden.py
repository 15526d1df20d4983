"""Generates tests/golden/*.npz from the numpy oracle (oracle/accfft_oracle.py).

The reference (AccFFT) holds no golden vectors and cannot be built in this image (no MPI, no FFTW), so
these fixtures are outputs of the ORACLE on seeded inputs -- they freeze the oracle's numbers so that a
later change of numpy / pocketfft or of the oracle itself cannot silently move the parity target.  The
oracle in turn is pinned to FFTW's own numbers and to the DFT definition by tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import accfft_oracle as orc  # noqa: E402

R2C = [((16, 12, 20), 101), ((8, 8, 8), 102), ((12, 10, 6), 103), ((32, 32, 32), 104), ((5, 6, 10), 105)]
C2C = [((16, 12, 20), 201), ((6, 10, 5), 202), ((24, 16, 12), 203)]

for n, seed in R2C:
    x = orc.synthetic_real(n, seed)
    np.savez_compressed(os.path.join(HERE, "r2c_%dx%dx%d.npz" % n), kind="r2c", n=np.array(n), seed=seed, x=x,
                        X=orc.forward_r2c(x))
for n, seed in C2C:
    z = orc.synthetic_complex(n, seed)
    np.savez_compressed(os.path.join(HERE, "c2c_%dx%dx%d.npz" % n), kind="c2c", n=np.array(n), seed=seed, x=z,
                        X=orc.forward_c2c(z))
n = (16, 20, 24)
x = orc.synthetic_real(n, 301)
np.savez_compressed(os.path.join(HERE, "ops_%dx%dx%d.npz" % n), kind="ops", n=np.array(n), seed=301, x=x,
                    laplace=orc.laplace(x), biharmonic=orc.biharmonic(x), gx=orc.grad(x, 0), gy=orc.grad(x, 1),
                    gz=orc.grad(x, 2))
print("wrote", len(R2C) + len(C2C) + 1, "fixtures to", HERE)
