"""CPU oracle for the AccFFT GPU hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product (``accfft_b200`` and
``libaccfft_b200.so``) never does and has no CPU fallback.

What this restates (citations are into /root/reference, which is NOT read at run time):

* pencil decomposition         src/accfft_common.cpp:49-91   (dfft_get_local_size_t)
* r2c / c2c size queries       src/accfft_common.cpp:98-192, :203-256
* Cartesian rank <-> coords    src/accfft_common.cpp:275-305 (MPI_Cart_create, row-major)
* MPI_Dims_create fallback     src/accfft_common.cpp:281-289
* transform definition         src/accfft.cpp:1352-1453 (forward), :1455-1551 (inverse):
                               unnormalised DFT, sign -1 forward / +1 inverse, r2c keeps k2 in [0, N2/2]
* staged algorithm             src/accfft_gpu.cpp:472-539 (Z fft, exchange 1, Y fft, exchange 2, X fft)
* exchange pack / unpack       src/transpose.cpp:2800-2815 (pack), :2894-2911 (unpack), ctor :150-330
* spectral operators           src/operators.txx:24-217 == src/operators_cuda.cu:15-205
* analytic fixtures            steps/step1/step1.cpp:16-26, steps/step3/step3_gpu.cpp:147-185,
                               steps/step5/step5_gpu.cpp:20-133

The arithmetic of the reference lives in FFTW3 (un-vendored, un-pinned dependency,
cmake/FindFFTW.cmake:44-71) which is absent from this image, as is MPI.  The oracle therefore
restates FFTW's *published definition* of the unnormalised DFT with numpy's pocketfft in float64.

PARITY PINNING: the reference holds no golden vectors and none of its tests inspect forward
spectra.  The oracle is pinned by (a) the reference's own analytic acceptance tests (step1/2/3 round
trip < 1e-10, step5 closed-form grad/laplace/biharmonic < 1e-10) which ``tests/test_oracle.py``
replays against this module, (b) FFTW's own double-precision 1-D reference vectors that ship inside
scipy (``scipy/fftpack/tests/fftw_double_ref.npz``), and (c) an O(N^2) long-double DFT written from
the definition.  Forward-spectrum *values* remain "parity unpinned" in the strict sense that no
reference-run output exists to compare with.  Independently of this module, the reference's own GPU
step programs are compiled unmodified against the product (oracle/Makefile -> oracle/_ref/steps) and must print their
own "Results are CORRECT!" verdict on the GPU (tests/test_reference_steps.py).
"""
from __future__ import annotations

import math
import os
from typing import List, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# decomposition
# --------------------------------------------------------------------------------------


def dims_create(nprocs: int) -> Tuple[int, int]:
    """MPI_Dims_create(nprocs, 2, {0,0}): balanced factors, non-increasing.

    src/accfft_common.cpp:281-289 falls back to this when c_dims[0]*c_dims[1] != nprocs
    (2 -> 2x1, 4 -> 2x2, 8 -> 4x2, 6 -> 3x2).
    """
    best = (nprocs, 1)
    for a in range(1, int(math.isqrt(nprocs)) + 1):
        if nprocs % a == 0:
            best = (nprocs // a, a)
    return best


def rank_to_coords(rank: int, np_: Sequence[int]) -> Tuple[int, int]:
    """Row-major Cartesian mapping used by MPI_Cart_create (src/accfft_common.cpp:298)."""
    return rank // np_[1], rank % np_[1]


def coords_to_rank(coords: Sequence[int], np_: Sequence[int]) -> int:
    return coords[0] * np_[1] + coords[1]


def block(N: int, P: int, c: int) -> Tuple[int, int]:
    """(size, start) of ceil-block c of N over P.  src/accfft_common.cpp:58-74."""
    size = int(math.ceil(N / float(P)))
    start = size * c
    if N - size * c < size:
        size = N - size * c
        size = size if size > 0 else 0
        start = N - size
    return size, start


def dfft_get_local_size(N0, N1, tuple_, np_, coords, sizeof_T):
    """src/accfft_common.cpp:49-91.  Returns (isize, istart, alloc_local_bytes)."""
    s0, st0 = block(N0, np_[0], coords[0])
    s1, st1 = block(N1, np_[1], coords[1])
    isize = [s0, s1, tuple_]
    istart = [st0, st1, 0]
    return isize, istart, s0 * s1 * tuple_ * sizeof_T


def _swap(a, i, j):
    a[i], a[j] = a[j], a[i]


def local_size_dft_r2c(n, np_, coords, sizeof_T=8):
    """src/accfft_common.cpp:98-192.  Returns dict with isize/istart/osize/ostart/alloc_max
    plus every intermediate stage size the plan uses (src/accfft_gpu.cpp:127-168)."""
    n0, n1, n2 = n
    n_tuples = (n2 // 2 + 1) * 2
    alloc_max = 0
    osize_0, ostart_0, a = dfft_get_local_size(n0, n1, n_tuples, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, a)
    osize_1, ostart_1, a = dfft_get_local_size(n0, n_tuples // 2, n1, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, a * 2)
    osize_2, ostart_2, a = dfft_get_local_size(n1, n_tuples // 2, n0, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, a * 2)
    _swap(osize_1, 1, 2); _swap(ostart_1, 1, 2)
    _swap(ostart_2, 1, 2); _swap(ostart_2, 0, 1)
    _swap(osize_2, 1, 2); _swap(osize_2, 0, 1)
    isize, istart, _ = dfft_get_local_size(n0, n1, n2, np_, coords, sizeof_T)
    # y-only
    osize_y, ostart_y, a = dfft_get_local_size(n0, n2, n1, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, a)
    osize_yi, ostart_yi, a = dfft_get_local_size(n0, n2, n1 // 2 + 1, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, 2 * a)
    _swap(osize_y, 1, 2); _swap(ostart_y, 1, 2)
    _swap(osize_yi, 1, 2); _swap(ostart_yi, 1, 2)
    # x-only: (N1/P1 * N2) split completely over P0
    osize_x, _, _ = dfft_get_local_size(isize[1] * n2, n2, n0, np_, coords, sizeof_T)
    osize_x = [osize_x[0], osize_x[2], 1]          # ((N1/P1*N2)/P0, N0, 1)
    alloc_max = max(alloc_max, osize_x[0] * osize_x[1] * osize_x[2] * sizeof_T)
    osize_xi, _, _ = dfft_get_local_size(isize[1] * n2, n2, n0 // 2 + 1, np_, coords, sizeof_T)
    osize_xi = [osize_xi[0], osize_xi[2], 1]
    alloc_max = max(alloc_max, 2 * osize_xi[0] * osize_xi[1] * osize_xi[2] * sizeof_T)
    # the plan stores osize_x/xi as (N0, M/P0, 1)  (src/accfft_gpu.cpp:146-158)
    return dict(isize=isize, istart=istart, osize=list(osize_2), ostart=list(ostart_2),
                osize_0=[osize_0[0], osize_0[1], n_tuples // 2], ostart_0=ostart_0,
                osize_1=osize_1, ostart_1=ostart_1, osize_2=osize_2, ostart_2=ostart_2,
                osize_y=osize_y, ostart_y=ostart_y, osize_yi=osize_yi,
                osize_x=[osize_x[1], osize_x[0], 1], osize_xi=[osize_xi[1], osize_xi[0], 1],
                alloc_max=alloc_max)


def local_size_dft_c2c(n, np_, coords, sizeof_T=8):
    """src/accfft_common.cpp:203-256."""
    n0, n1, n2 = n
    alloc_max = 0
    osize_0, ostart_0, a = dfft_get_local_size(n0, n1, n2, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, a)
    osize_1, ostart_1, a = dfft_get_local_size(n0, n2, n1, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, a)
    osize_2, ostart_2, a = dfft_get_local_size(n1, n2, n0, np_, coords, sizeof_T)
    alloc_max = max(alloc_max, a)
    alloc_max *= 2
    _swap(osize_1, 1, 2); _swap(ostart_1, 1, 2)
    _swap(ostart_2, 1, 2); _swap(ostart_2, 0, 1)
    _swap(osize_2, 1, 2); _swap(osize_2, 0, 1)
    isize, istart, _ = dfft_get_local_size(n0, n1, n2, np_, coords, sizeof_T)
    return dict(isize=isize, istart=istart, osize=list(osize_2), ostart=list(ostart_2),
                osize_0=osize_0, ostart_0=ostart_0, osize_1=osize_1, ostart_1=ostart_1,
                osize_2=osize_2, ostart_2=ostart_2, alloc_max=alloc_max)


# --------------------------------------------------------------------------------------
# global transforms (the definition FFTW implements) and per-rank views
# --------------------------------------------------------------------------------------


def forward_r2c(x: np.ndarray) -> np.ndarray:
    """Unnormalised forward r2c, all three axes: X[k0,k1,k2], k2 in [0, N2/2]."""
    return np.fft.rfftn(np.asarray(x, dtype=np.float64), axes=(0, 1, 2))


def inverse_c2r(X: np.ndarray, n: Sequence[int]) -> np.ndarray:
    """Unnormalised inverse c2r: c2r(r2c(x)) == N0*N1*N2 * x  (src/accfft_gpu.cpp:540-604)."""
    N = int(n[0]) * int(n[1]) * int(n[2])
    return np.fft.irfftn(np.asarray(X, dtype=np.complex128), s=tuple(n), axes=(0, 1, 2)) * N


def forward_c2c(x: np.ndarray) -> np.ndarray:
    return np.fft.fftn(np.asarray(x, dtype=np.complex128), axes=(0, 1, 2))


def inverse_c2c(X: np.ndarray) -> np.ndarray:
    X = np.asarray(X, dtype=np.complex128)
    return np.fft.ifftn(X, axes=(0, 1, 2)) * X.size


def _axes_of(xyz: int):
    """bit 2 = Z (axis 2), bit 1 = Y (axis 1), bit 0 = X (axis 0): the std::bitset<3> xyz of
    accfft_execute_gpu / accfft_execute_c2c_gpu (src/accfft_gpu.cpp:477,499,527; :975,1000,1029)."""
    return [ax for ax, bit in ((0, 1), (1, 2), (2, 4)) if xyz & bit]


def forward_partial(x: np.ndarray, xyz: int, r2c: bool) -> np.ndarray:
    """Partial-axis forward transform: only the axes selected in xyz are transformed, the data still ends in the
    output layout (the reference always runs its transposes).  r2c needs the Z bit (src/accfft_gpu.cpp:477-488)."""
    if r2c:
        assert xyz & 4
        X = np.fft.rfft(np.asarray(x, dtype=np.float64), axis=2)
        axes = [a for a in _axes_of(xyz) if a != 2]
        return np.fft.fftn(X, axes=axes) if axes else X
    axes = _axes_of(xyz)
    z = np.asarray(x, dtype=np.complex128)
    return np.fft.fftn(z, axes=axes) if axes else z.copy()


def inverse_partial(X: np.ndarray, n, xyz: int, r2c: bool) -> np.ndarray:
    """Unnormalised partial-axis inverse (mirror of forward_partial)."""
    if r2c:
        assert xyz & 4
        axes = [a for a in _axes_of(xyz) if a != 2]
        Y = np.fft.ifftn(X, axes=axes) * np.prod([n[a] for a in axes]) if axes else X
        return np.fft.irfft(Y, n=n[2], axis=2) * n[2]
    axes = _axes_of(xyz)
    return np.fft.ifftn(X, axes=axes) * np.prod([n[a] for a in axes]) if axes else np.array(X, copy=True)


def local_input(x: np.ndarray, n, np_, coords) -> np.ndarray:
    """Rank (c0,c1)'s input block x[istart0:+isize0, istart1:+isize1, :]  (steps/step1/step1.cpp:162-170)."""
    s0, st0 = block(n[0], np_[0], coords[0])
    s1, st1 = block(n[1], np_[1], coords[1])
    return np.ascontiguousarray(x[st0:st0 + s0, st1:st1 + s1, :])


def local_output(X: np.ndarray, n, np_, coords, r2c=True) -> np.ndarray:
    """Rank (c0,c1)'s output block X[:, ostart1:+osize1, ostart2:+osize2]."""
    n2c = n[2] // 2 + 1 if r2c else n[2]
    s1, st1 = block(n[1], np_[0], coords[0])
    s2, st2 = block(n2c, np_[1], coords[1])
    return np.ascontiguousarray(X[:, st1:st1 + s1, st2:st2 + s2])


def assemble_output(blocks: List[np.ndarray], n, np_, r2c=True) -> np.ndarray:
    """Inverse of local_output over all ranks (rank order)."""
    n2c = n[2] // 2 + 1 if r2c else n[2]
    X = np.zeros((n[0], n[1], n2c), dtype=blocks[0].dtype)
    for r, b in enumerate(blocks):
        c = rank_to_coords(r, np_)
        s1, st1 = block(n[1], np_[0], c[0])
        s2, st2 = block(n2c, np_[1], c[1])
        X[:, st1:st1 + s1, st2:st2 + s2] = b.reshape(n[0], s1, s2)
    return X


def assemble_input(blocks: List[np.ndarray], n, np_) -> np.ndarray:
    x = np.zeros(tuple(n), dtype=blocks[0].dtype)
    for r, b in enumerate(blocks):
        c = rank_to_coords(r, np_)
        s0, st0 = block(n[0], np_[0], c[0])
        s1, st1 = block(n[1], np_[1], c[1])
        x[st0:st0 + s0, st1:st1 + s1, :] = b.reshape(s0, s1, n[2])
    return x


# --------------------------------------------------------------------------------------
# single-axis real transforms: accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t (src/accfft_gpu.cpp:1212-1604)
# --------------------------------------------------------------------------------------
def axis_forward(x: np.ndarray, axis: int) -> np.ndarray:
    """Global result of the single-axis r2c: unnormalised real DFT along `axis` only (0 = x, 1 = y, 2 = z),
    keeping k in [0, N_axis/2] (cufftExecD2Z of fplan_x / fplan_y / fplan_0, src/accfft_gpu.cpp:1234, :1370, :1524)."""
    return np.fft.rfft(np.asarray(x, np.float64), axis=axis)


def axis_inverse(X: np.ndarray, n, axis: int) -> np.ndarray:
    """Global result of the single-axis c2r: unnormalised (c2r(r2c(x)) = N_axis * x)."""
    return np.fft.irfft(X, n=n[axis], axis=axis) * n[axis]


def axis_local_spectrum(Xa: np.ndarray, n, np_, coords, axis: int) -> np.ndarray:
    """Rank (c0,c1)'s block of the single-axis spectrum Xa = axis_forward(x, axis), in the reference's layouts:
      z: [isize0][isize1][N2/2+1]                       (osize_0,  src/accfft_common.cpp:110-112)
      y: [isize0][N1/2+1][block of N2 over P1]          (osize_yi, :150-162: N2 is split over the row group)
      x: [N0/2+1][block of the flattened local plane isize1*N2 over P0]   (osize_xi, :164-188)"""
    s0, st0 = block(n[0], np_[0], coords[0])
    s1, st1 = block(n[1], np_[1], coords[1])
    if axis == 2:
        return np.ascontiguousarray(Xa[st0:st0 + s0, st1:st1 + s1, :])
    if axis == 1:
        s2, st2 = block(n[2], np_[1], coords[1])
        return np.ascontiguousarray(Xa[st0:st0 + s0, :, st2:st2 + s2])
    M = s1 * n[2]
    sm, stm = block(M, np_[0], coords[0])
    plane = Xa[:, st1:st1 + s1, :].reshape(Xa.shape[0], M)
    return np.ascontiguousarray(plane[:, stm:stm + sm])


# --------------------------------------------------------------------------------------
# staged (distributed) algorithm with the reference's own pack / unpack index formulas
# --------------------------------------------------------------------------------------


def _tplan_exchange(blocks: List[np.ndarray], N0: int, N1: int, howmany: int) -> List[np.ndarray]:
    """Forward "transposed-out" exchange among len(blocks) ranks, complex elements.

    Rank p holds, for h in howmany, a [local_n0(p)][N1] array; afterwards rank q holds, for h,
    [N0][local_n1(q)].  Pack:  src/transpose.cpp:2800-2815;  unpack: :2894-2911;
    counts (clean rule): send p->q = howmany*local_n0(p)*local_n1(q).
    """
    P = len(blocks)
    ln0 = [block(N0, P, p) for p in range(P)]
    ln1 = [block(N1, P, p) for p in range(P)]
    send = [[None] * P for _ in range(P)]
    for p in range(P):
        d = blocks[p].reshape(howmany, ln0[p][0], N1)
        for q in range(P):
            send[p][q] = d[:, :, ln1[q][1]:ln1[q][1] + ln1[q][0]].copy()   # [h][i][j-run]
    out = []
    for q in range(P):
        o = np.zeros((howmany, N0, ln1[q][0]), dtype=blocks[0].dtype)
        for p in range(P):
            o[:, ln0[p][1]:ln0[p][1] + ln0[p][0], :] = send[p][q]
        out.append(o)
    return out


def staged_forward(x_or_z: np.ndarray, n, np_, r2c=True) -> List[np.ndarray]:
    """Run the reference's distributed algorithm (src/accfft_gpu.cpp:472-539) rank by rank in one
    process and return each rank's final [N0][osize1][osize2] block (rank order)."""
    P0, P1 = np_
    n0, n1, n2 = n
    n2c = n2 // 2 + 1 if r2c else n2
    nr = P0 * P1
    # stage Z
    st = []
    for r in range(nr):
        c = rank_to_coords(r, np_)
        loc = local_input(x_or_z, n, np_, c)
        st.append(np.fft.rfft(loc, axis=2) if r2c else np.fft.fft(loc, axis=2))
    # exchange 1 in each row group (same c0): per i0 slab  [n1l][N2c] -> [N1][k2l]
    st1 = [None] * nr
    for c0 in range(P0):
        grp = [coords_to_rank((c0, c1), np_) for c1 in range(P1)]
        n0l = block(n0, P0, c0)[0]
        res = _tplan_exchange([st[r] for r in grp], n1, n2c, n0l)
        for c1, r in enumerate(grp):
            st1[r] = res[c1]                       # [n0l][N1][k2l]
    # stage Y
    st1 = [np.fft.fft(a, axis=1) for a in st1]
    # exchange 2 in each column group (same c1): [n0l][N1*k2l] -> [N0][n1l'*k2l], n_tuples = k2l
    st2 = [None] * nr
    for c1 in range(P1):
        grp = [coords_to_rank((c0, c1), np_) for c0 in range(P0)]
        k2l = block(n2c, P1, c1)[0]
        blocks = []
        for r in grp:
            blocks.append(st1[r].reshape(1, st1[r].shape[0], n1, k2l))
        # treat k2l as the tuple dimension: do the exchange per k2 column
        ln0 = [block(n0, P0, p) for p in range(P0)]
        ln1 = [block(n1, P0, p) for p in range(P0)]
        for q, r in enumerate(grp):
            o = np.zeros((n0, ln1[q][0], k2l), dtype=np.complex128)
            for p in range(P0):
                o[ln0[p][1]:ln0[p][1] + ln0[p][0]] = st1[grp[p]][:, ln1[q][1]:ln1[q][1] + ln1[q][0], :]
            st2[r] = o
    # stage X
    return [np.fft.fft(a, axis=0) for a in st2]


# --------------------------------------------------------------------------------------
# spectral operators  (src/operators.txx:24-217, src/operators_cuda.cu:15-205)
# --------------------------------------------------------------------------------------


def wave_numbers(N: int, count: int | None = None) -> np.ndarray:
    """k = X > N/2 ? X-N : X, and k = 0 at X == N/2 (Nyquist filtered).  Integer division as in C."""
    X = np.arange(N if count is None else count)
    k = np.where(X > N // 2, X - N, X)
    k = np.where(X == N // 2, 0, k)
    return k.astype(np.float64)


def _khat(n):
    n2c = n[2] // 2 + 1
    kx = wave_numbers(n[0])[:, None, None]
    ky = wave_numbers(n[1])[None, :, None]
    kz = wave_numbers(n[2], n2c)[None, None, :]
    return kx, ky, kz


def grad(x: np.ndarray, axis: int) -> np.ndarray:
    """accfft_grad component: ifft( i k_axis * fft(x) ) / N  (src/operators_gpu.txx:172-278).
    The fast path transforms along one axis only; mathematically identical."""
    n = x.shape
    k = wave_numbers(n[axis])
    # the reference transforms along the axis with r2c/c2r, so it only sees X <= N/2; c2r implies the mirrored
    # half, hence the pair +-floor(N/2) is zeroed as a pair (matters only for odd N)
    k[n[axis] - n[axis] // 2 if n[axis] > 1 else 0] = 0.0
    shape = [1, 1, 1]
    shape[axis] = n[axis]
    X = np.fft.fft(np.asarray(x, np.float64), axis=axis)
    X = X * (1j * k.reshape(shape))
    return np.fft.ifft(X, axis=axis).real


def grad_slow(x: np.ndarray, axis: int) -> np.ndarray:
    """accfft_grad_slow_t component (src/operators.txx / src/operators_gpu.txx:72-170): full 3-D r2c, multiply by
    i k_axis / N with ONLY X == N/2 zeroed (src/operators_cuda.cu:34-37), full 3-D c2r.  Identical to grad() for even
    N; for odd N along x or y the mirror mode N - N/2 survives here (the fast scheme's real transform drops the pair)."""
    n = x.shape
    shape = [1, 1, 1]
    shape[axis] = n[axis] if axis < 2 else n[2] // 2 + 1
    k = wave_numbers(n[axis], shape[axis]).reshape(shape)
    X = forward_r2c(x) * (1j * k)
    return inverse_c2r(X, n) / (n[0] * n[1] * n[2])


def divergence_slow(ax: np.ndarray, ay: np.ndarray, az: np.ndarray) -> np.ndarray:
    """accfft_divergence_slow_t: sum of the three grad_slow-style derivatives (one inverse transform of the summed spectra)."""
    return grad_slow(ax, 0) + grad_slow(ay, 1) + grad_slow(az, 2)


def laplace(x: np.ndarray) -> np.ndarray:
    n = x.shape
    kx, ky, kz = _khat(n)
    X = forward_r2c(x) * (-(kx * kx + ky * ky + kz * kz))
    return inverse_c2r(X, n) / (n[0] * n[1] * n[2])


def biharmonic(x: np.ndarray) -> np.ndarray:
    n = x.shape
    kx, ky, kz = _khat(n)
    s = kx * kx + ky * ky + kz * kz
    X = forward_r2c(x) * (s * s)
    return inverse_c2r(X, n) / (n[0] * n[1] * n[2])


def _k2(n):
    kx, ky, kz = _khat(n)
    return kx * kx + ky * ky + kz * kz


def inv_laplace(x: np.ndarray) -> np.ndarray:
    """accfft_inv_laplace (CPU-only in the reference): F^-1[F[x] / -(|k|^2)], 0 where |k|^2 == 0 with the Nyquist-zeroed
    wave numbers  (src/operators.txx:275-333: wave = -wx^2-wy^2-wz^2; wave == 0 ? 0 : 1/wave)."""
    n = x.shape
    k2 = _k2(n)
    w = np.zeros_like(k2)
    np.divide(-1.0, k2, out=w, where=k2 != 0)
    return np.fft.irfftn(np.fft.rfftn(np.asarray(x, np.float64)) * w, s=n, axes=(0, 1, 2))


def _k2_plain(n):
    """|k|^2 with the plain signed wave numbers (X > N/2 -> X - N; X == N/2 stays N/2): the reference's inv_biharmonic has
    the Nyquist rule commented out (src/operators.txx:362-375)."""
    def k(N, count=None):
        X = np.arange(N if count is None else count)
        return np.where(X > N // 2, X - N, X).astype(np.float64)
    kx = k(n[0])[:, None, None]
    ky = k(n[1])[None, :, None]
    kz = k(n[2], n[2] // 2 + 1)[None, None, :]
    return kx * kx + ky * ky + kz * kz


def inv_biharmonic(x: np.ndarray) -> np.ndarray:
    """accfft_inv_biharmonic (src/operators.txx:335-392), as the reference computes it (pinned by tests/golden/ref_*.npz):
    F^-1[F[x] / |k|^4] with the PLAIN wave numbers (no Nyquist zeroing, :362-375) and the zero mode passed through
    unchanged (`if (wave == 0) wave = 1`, :379-381) -- unlike accfft_inv_laplace, which zeroes both."""
    n = x.shape
    k2 = _k2_plain(n)
    w = np.ones_like(k2)
    np.divide(1.0, k2 * k2, out=w, where=k2 != 0)
    return np.fft.irfftn(np.fft.rfftn(np.asarray(x, np.float64)) * w, s=n, axes=(0, 1, 2))


def symbol_solve(x: np.ndarray, alpha: float, beta: float, inverse: bool) -> np.ndarray:
    """F^-1[(alpha + beta |k|^2)^(+-1) F[x]]; inverse: modes with a zero denominator -> 0.  Restates the Poisson filter
    (steps/step4/kernels.cu:86-104, method 1: alpha = 0, beta = -4 pi^2, zero mode -> 0; note that kernel uses the plain
    signed wave numbers without the Nyquist rule -- the operators' rule is used here, they differ only on X == N/2) and
    the explicit heat update u + dt * laplace(u) of steps/step6/step6.cpp:150-165 (alpha = 1, beta = -dt)."""
    n = x.shape
    sym = alpha + beta * _k2(n)
    if inverse:
        w = np.zeros_like(sym)
        np.divide(1.0, sym, out=w, where=sym != 0)
    else:
        w = sym
    return np.fft.irfftn(np.fft.rfftn(np.asarray(x, np.float64)) * w, s=n, axes=(0, 1, 2))


def divergence(ax: np.ndarray, ay: np.ndarray, az: np.ndarray) -> np.ndarray:
    """src/operators_gpu.txx:448-553: d/dx ax + d/dy ay + d/dz az."""
    return grad(ax, 0) + grad(ay, 1) + grad(az, 2)


# --------------------------------------------------------------------------------------
# analytic fixtures of the reference's step programs
# --------------------------------------------------------------------------------------


def grid(n):
    X = 2 * math.pi / n[0] * np.arange(n[0])[:, None, None]
    Y = 2 * math.pi / n[1] * np.arange(n[1])[None, :, None]
    Z = 2 * math.pi / n[2] * np.arange(n[2])[None, None, :]
    return X, Y, Z


def step1_testcase(n) -> np.ndarray:
    """steps/step1/step1.cpp:16-26: exp(-4 * |x - pi|^2)."""
    X, Y, Z = grid(n)
    pi = math.pi
    return np.exp(-4.0 * ((X - pi) ** 2 + (Y - pi) ** 2 + (Z - pi) ** 2))


def step5_testcase(n, sigma=32.0):
    """steps/step5/step5_gpu.cpp:20-133: f, grad f (3), laplace f, biharmonic f in closed form."""
    X, Y, Z = grid(n)
    pi = math.pi
    Xp, Yp, Zp = X - pi, Y - pi, Z - pi
    g = np.exp(-sigma * (Xp ** 2 + Yp ** 2 + Zp ** 2))
    sx, sy, sz = np.sin(X), np.sin(Y), np.sin(Z)
    f = sx * sy * sz + g
    gx = np.cos(X) * sy * sz + (-2 * sigma * Xp) * g
    gy = sx * np.cos(Y) * sz + (-2 * sigma * Yp) * g
    gz = sx * sy * np.cos(Z) + (-2 * sigma * Zp) * g
    lap = (4 * sigma ** 2 * (Xp ** 2 + Yp ** 2 + Zp ** 2) - 6 * sigma) * g - 3 * sx * sy * sz
    s2, s3, s4 = sigma ** 2, sigma ** 3, sigma ** 4

    def d4(a):
        return 12 * s2 * g - 48 * s3 * a * a * g + 16 * s4 * a ** 4 * g

    def d22(a, b):
        return 4 * s2 * g - 8 * s3 * b * b * g - 8 * s3 * a * a * g + 16 * s4 * a * a * b * b * g

    bih = 9 * sx * sy * sz + (d4(Xp) + d4(Yp) + d4(Zp) + 2 * d22(Xp, Yp) + 2 * d22(Xp, Zp) + 2 * d22(Yp, Zp))
    return dict(f=f, gx=gx, gy=gy, gz=gz, laplace=lap, biharmonic=bih)


def synthetic_real(n, seed: int, dist="uniform") -> np.ndarray:
    """x ~ U(-1,1) (or N(0,1)), numpy default_rng(seed), generated globally."""
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        return rng.uniform(-1.0, 1.0, size=tuple(n))
    return rng.standard_normal(size=tuple(n))


def synthetic_complex(n, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, size=tuple(n)) + 1j * rng.uniform(-1.0, 1.0, size=tuple(n))


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """relative L2 error |a-b|_2 / |b|_2 (north_star tolerance: 1e-12 double, 1e-5 single)."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    d = np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128))
    nb = np.linalg.norm(b.astype(np.complex128))
    return float(d / nb) if nb > 0 else float(d)


# --------------------------------------------------------------------------------------
# definition-level DFT (long double, O(N^2)) used to pin pocketfft itself
# --------------------------------------------------------------------------------------


def dft_definition(x: np.ndarray, sign=-1) -> np.ndarray:
    """X[k] = sum_n x[n] exp(sign * 2 pi i n k / N) in long double, straight from the definition."""
    x = np.asarray(x)
    N = x.shape[0]
    nk = (np.arange(N)[:, None] * np.arange(N)[None, :]) % N
    ang = (2 * np.pi * nk.astype(np.longdouble)) / np.longdouble(N)
    re = np.cos(ang)
    im = sign * np.sin(ang)
    xr = x.real.astype(np.longdouble)
    xi = x.imag.astype(np.longdouble) if np.iscomplexobj(x) else np.zeros(N, np.longdouble)
    Xr = re @ xr - im @ xi
    Xi = re @ xi + im @ xr
    return (Xr + 1j * Xi).astype(np.complex128)


# --------------------------------------------------------------------------------------
# CPU baseline timing stand-in: pocketfft, all host cores
# --------------------------------------------------------------------------------------


def cpu_input(n, dtype=np.float64, c2c=False):
    rng = np.random.default_rng(1)
    if c2c:
        return (rng.standard_normal(tuple(n)) + 1j * rng.standard_normal(tuple(n))).astype(
            np.complex128 if dtype == np.float64 else np.complex64)
    return rng.standard_normal(tuple(n)).astype(dtype)


def cpu_roundtrip_seconds(n, dtype=np.float64, repeats=1, workers=None, c2c=False, x=None):
    """Time one r2c+c2r (or c2c fwd+inv) round trip with scipy.fft (pocketfft, `workers` threads).
    Stand-in for the reference's FFTW+MPI CPU path, which cannot be built in this image.
    Only the two transforms are timed (input generation is outside).  Returns (best seconds, workers)."""
    import time
    import scipy.fft as sfft
    workers = workers or (os.cpu_count() or 1)
    if x is None:
        x = cpu_input(n, dtype, c2c)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        if c2c:
            X = sfft.fftn(x, workers=workers)
            y = sfft.ifftn(X, workers=workers)
        else:
            X = sfft.rfftn(x, workers=workers)
            y = sfft.irfftn(X, s=tuple(n), workers=workers)
        best = min(best, time.perf_counter() - t0)
    del y
    return best, workers


def hashed_field(n, seed: int, c2c=False) -> np.ndarray:
    """Seeded test field that C++ can regenerate bit for bit (oracle/ref_driver.cpp: hashed()): splitmix64 of
    (seed, linear global index) mapped to [-1, 1).  Complex fields interleave re / im at indices 2g, 2g+1."""
    total = int(np.prod(n)) * (2 if c2c else 1)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.arange(total, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    v = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0
    if c2c:
        return (v[0::2] + 1j * v[1::2]).reshape(tuple(n))
    return v.reshape(tuple(n))
