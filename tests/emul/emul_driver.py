"""ctypes driver for tests/emul/_build/libemul.so (CPU emulator of the stage kernels; test infra only)."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libemul.so")
NPARTS = 8
MAIN_SOURCES = [os.path.join(HERE, "emul.cpp"), os.path.join(ROOT, "accfft_b200", "csrc", "plan.cpp"), os.path.join(HERE, "fused_sim.cpp")]
PART_SOURCE = os.path.join(HERE, "emul_part.cpp")
ALT_SOURCE = os.path.join(HERE, "emul_alt.cpp")   # kind- / precision-specific schedules of the contiguous kernels
SOURCES = MAIN_SOURCES + [PART_SOURCE, ALT_SOURCE, os.path.join(HERE, "emul_fiber.hpp")]
HEADERS = [os.path.join(ROOT, "accfft_b200", "csrc", f)
           for f in ("fft_engine.cuh", "stage.cuh", "stage_real.cuh", "plan.hpp", "launch.hpp", "decomp.hpp", "kernel_cfg.hpp", "schedules.def", "fused_sched.hpp")]

sys.path.insert(0, ROOT)
from oracle import accfft_oracle as orc  # noqa: E402


def build(force=False):
    """g++ only (no CUDA).  The kernel-body instantiations are split over NPARTS translation units (emul_part.cpp with
    -DEMU_PART=k) that compile in parallel: ~1 minute instead of ~5 for one unit."""
    bdir = os.path.dirname(LIB)
    os.makedirs(bdir, exist_ok=True)
    if not force and os.path.exists(LIB):
        m = os.path.getmtime(LIB)
        if all(os.path.getmtime(p) <= m for p in SOURCES + HEADERS):
            return LIB
    flags = ["-O1", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-DEMU_NPARTS=%d" % NPARTS,
             "-I", os.path.join(ROOT, "accfft_b200", "csrc")]
    jobs = []
    objs = []
    for k in range(NPARTS):
        o = os.path.join(bdir, "part%d.o" % k)
        objs.append(o)
        jobs.append(subprocess.Popen(["g++", *flags, "-DEMU_PART=%d" % k, "-c", PART_SOURCE, "-o", o]))
    for kind in (1, 2, 3):
        for prec in (0, 1):
            o = os.path.join(bdir, "alt_k%d_p%d.o" % (kind, prec))
            objs.append(o)
            jobs.append(subprocess.Popen(["g++", *flags, "-DEMU_ALT_KIND=%d" % kind, "-DEMU_ALT_PREC=%d" % prec, "-c", ALT_SOURCE, "-o", o]))
    for i, src in enumerate(MAIN_SOURCES):
        o = os.path.join(bdir, "main%d.o" % i)
        objs.append(o)
        jobs.append(subprocess.Popen(["g++", *flags, "-x", "c++", "-c", src, "-o", o]))
    failed = [j.args for j in jobs if j.wait() != 0]
    if failed:
        raise RuntimeError("emulator build failed: %r" % (failed[0],))
    subprocess.check_call(["g++", "-shared", "-o", LIB, *objs, "-lpthread"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_world_create.restype = ctypes.c_void_p
        _lib.emu_world_create.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int]
        _lib.emu_world_destroy.argtypes = [ctypes.c_void_p]
        _lib.emu_alloc_max.restype = ctypes.c_longlong
        _lib.emu_alloc_max.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib.emu_execute.restype = ctypes.c_int
        _lib.emu_execute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        _lib.emu_execute_xyz.restype = ctypes.c_int
        _lib.emu_execute_xyz.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                         ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        pp = ctypes.POINTER(ctypes.c_void_p)
        _lib.emu_alt_schedule_count.restype = ctypes.c_int
        _lib.emu_alt_launches.restype = ctypes.c_longlong
        _lib.emu_fused_protocol_sim.restype = ctypes.c_int
        _lib.emu_fused_protocol_sim.argtypes = [ctypes.c_int] * 9 + [ctypes.c_ulonglong, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        _lib.emu_execute_many.restype = ctypes.c_int
        _lib.emu_execute_many.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, pp, pp]
        _lib.emu_execute_grad.restype = ctypes.c_int
        _lib.emu_execute_grad.argtypes = [ctypes.c_void_p, pp, pp, pp, pp, ctypes.c_int]
        _lib.emu_execute_div.restype = ctypes.c_int
        _lib.emu_execute_div.argtypes = [ctypes.c_void_p, pp, pp, pp, pp]
        _lib.emu_execute_symbol.restype = ctypes.c_int
        _lib.emu_execute_symbol.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, pp, pp]
        _lib.emu_execute_axis.restype = ctypes.c_int
        _lib.emu_execute_axis.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, pp, pp]
        _lib.emu_execute_grad_axes.restype = ctypes.c_int
        _lib.emu_execute_grad_axes.argtypes = [ctypes.c_void_p, pp, pp, pp, pp, ctypes.c_int]
        _lib.emu_execute_div_axes.restype = ctypes.c_int
        _lib.emu_execute_div_axes.argtypes = [ctypes.c_void_p, pp, pp, pp, pp]
    return _lib


class World:
    """All ranks of a P0 x P1 grid in one process."""

    def __init__(self, n, np_, r2c=True, single=False):
        self.n, self.np_, self.r2c, self.single = tuple(n), tuple(np_), r2c, single
        arr = (ctypes.c_int * 3)(*n)
        self.h = lib().emu_world_create(arr, np_[0], np_[1], int(r2c), int(single))
        if not self.h:
            raise RuntimeError("emu_world_create failed (unsupported size?)")
        self.nprocs = np_[0] * np_[1]
        self.rdt = np.float32 if single else np.float64
        self.cdt = np.complex64 if single else np.complex128

    def close(self):
        if self.h:
            lib().emu_world_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def coords(self, r):
        return orc.rank_to_coords(r, self.np_)

    def _run(self, mode, ins, outs, padded=False, xyz=7):
        pin = (ctypes.c_void_p * self.nprocs)(*[a.ctypes.data for a in ins])
        pout = (ctypes.c_void_p * self.nprocs)(*[a.ctypes.data for a in outs])
        if xyz != 7:
            rc = lib().emu_execute_xyz(self.h, mode, pin, pout, xyz)
        else:
            rc = lib().emu_execute(self.h, mode, pin, pout, int(padded))
        if rc:
            raise RuntimeError(f"emu_execute rc={rc}")

    def alloc_out(self):
        """per-rank output buffers of alloc_max bytes"""
        return [np.zeros(lib().emu_alloc_max(self.h, r) // self.rdt().itemsize, dtype=self.rdt)
                for r in range(self.nprocs)]

    def forward(self, x_global, padded=False, xyz=7):
        n = self.n
        ins = []
        for r in range(self.nprocs):
            loc = orc.local_input(x_global, n, self.np_, self.coords(r))
            if self.r2c:
                loc = loc.astype(self.rdt)
                if padded:
                    n2c = n[2] // 2 + 1
                    buf = np.zeros(loc.shape[:2] + (2 * n2c,), dtype=self.rdt)
                    buf[:, :, :n[2]] = loc
                    loc = buf
            else:
                loc = loc.astype(self.cdt)
            ins.append(np.ascontiguousarray(loc))
        outs = self.alloc_out()
        self._run(0, ins, outs, padded, xyz)
        blocks = []
        n2c = n[2] // 2 + 1 if self.r2c else n[2]
        for r in range(self.nprocs):
            c = self.coords(r)
            s1 = orc.block(n[1], self.np_[0], c[0])[0]
            s2 = orc.block(n2c, self.np_[1], c[1])[0]
            cnt = n[0] * s1 * s2
            blocks.append(outs[r].view(self.cdt)[:cnt].reshape(n[0], s1, s2).copy())
        return blocks

    def inverse(self, X_global, padded=False, xyz=7):
        n = self.n
        ins = []
        for r in range(self.nprocs):
            loc = orc.local_output(X_global, n, self.np_, self.coords(r), r2c=self.r2c).astype(self.cdt)
            ins.append(np.ascontiguousarray(loc))
        outs = self.alloc_out()
        self._run(1, ins, outs, padded, xyz)
        blocks = []
        for r in range(self.nprocs):
            c = self.coords(r)
            s0 = orc.block(n[0], self.np_[0], c[0])[0]
            s1 = orc.block(n[1], self.np_[1], c[1])[0]
            if self.r2c:
                row = 2 * (n[2] // 2 + 1) if padded else n[2]
                blocks.append(outs[r][:s0 * s1 * row].reshape(s0, s1, row)[:, :, :n[2]].copy())
            else:
                blocks.append(outs[r].view(self.cdt)[:s0 * s1 * n[2]].reshape(s0, s1, n[2]).copy())
        return blocks

    def forward_many(self, xs_global):
        """batched forward transforms of several global fields (worlds made under EMU_BATCH >= len(xs_global))"""
        n, nf = self.n, len(xs_global)
        ins, outs = [], []
        for x_global in xs_global:
            ins.append([np.ascontiguousarray(orc.local_input(x_global, n, self.np_, self.coords(r)).astype(self.rdt if self.r2c else self.cdt))
                        for r in range(self.nprocs)])
            outs.append(self.alloc_out())
        pin = (ctypes.c_void_p * (nf * self.nprocs))(*[a.ctypes.data for f in ins for a in f])
        pout = (ctypes.c_void_p * (nf * self.nprocs))(*[a.ctypes.data for f in outs for a in f])
        rc = lib().emu_execute_many(self.h, 0, nf, pin, pout)
        if rc:
            raise RuntimeError(f"emu_execute_many rc={rc}")
        n2c = n[2] // 2 + 1 if self.r2c else n[2]
        res = []
        for f in range(nf):
            blocks = []
            for r in range(self.nprocs):
                c = self.coords(r)
                s1 = orc.block(n[1], self.np_[0], c[0])[0]
                s2 = orc.block(n2c, self.np_[1], c[1])[0]
                blocks.append(outs[f][r].view(self.cdt)[:n[0] * s1 * s2].reshape(n[0], s1, s2).copy())
            res.append(blocks)
        return res

    def inverse_many(self, Xs_global):
        n, nf = self.n, len(Xs_global)
        ins, outs = [], []
        for X_global in Xs_global:
            ins.append([np.ascontiguousarray(orc.local_output(X_global, n, self.np_, self.coords(r), r2c=self.r2c).astype(self.cdt))
                        for r in range(self.nprocs)])
            outs.append(self.alloc_out())
        pin = (ctypes.c_void_p * (nf * self.nprocs))(*[a.ctypes.data for f in ins for a in f])
        pout = (ctypes.c_void_p * (nf * self.nprocs))(*[a.ctypes.data for f in outs for a in f])
        rc = lib().emu_execute_many(self.h, 1, nf, pin, pout)
        if rc:
            raise RuntimeError(f"emu_execute_many rc={rc}")
        res = []
        for f in range(nf):
            blocks = []
            for r in range(self.nprocs):
                c = self.coords(r)
                s0 = orc.block(n[0], self.np_[0], c[0])[0]
                s1 = orc.block(n[1], self.np_[1], c[1])[0]
                if self.r2c:
                    blocks.append(outs[f][r][:s0 * s1 * n[2]].reshape(s0, s1, n[2]).copy())
                else:
                    blocks.append(outs[f][r].view(self.cdt)[:s0 * s1 * n[2]].reshape(s0, s1, n[2]).copy())
            res.append(blocks)
        return res

    def spectral(self, x_global, op):
        n = self.n
        ins = [np.ascontiguousarray(orc.local_input(x_global, n, self.np_, self.coords(r)).astype(self.rdt))
               for r in range(self.nprocs)]
        outs = self.alloc_out()
        self._run({"laplace": 2, "biharmonic": 3}[op], ins, outs)
        blocks = []
        for r in range(self.nprocs):
            c = self.coords(r)
            s0 = orc.block(n[0], self.np_[0], c[0])[0]
            s1 = orc.block(n[1], self.np_[1], c[1])[0]
            blocks.append(outs[r][:s0 * s1 * n[2]].reshape(s0, s1, n[2]).copy())
        return blocks


def fused_protocol_sim(nA, nBz, PCZ, tpa, nctas, cp=0, lag=0, ns=0, zfirst=True, seed=1, flags=0):
    """-> (rc, {"cp", "lag", "ns", "ring_planes", "chunks"}): a randomised CPU run of the fused kernel's scheduling protocol on the
    real work-list code (tests/emul/fused_sim.cpp); rc 0 = every dependency held, every tile ran once, no deadlock"""
    shape = (ctypes.c_int * 5)()
    rc = lib().emu_fused_protocol_sim(nA, nBz, PCZ, tpa, nctas, cp, lag, ns, int(zfirst), seed, flags, shape)
    return rc, dict(zip(("cp", "lag", "ns", "ring_planes", "chunks"), [int(v) for v in shape]))


def _local_real(w, x_global):
    return [np.ascontiguousarray(orc.local_input(x_global, w.n, w.np_, w.coords(r)).astype(w.rdt))
            for r in range(w.nprocs)]


def _ptrs(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _blocks_real(w, outs):
    n = w.n
    blocks = []
    for r in range(w.nprocs):
        c = w.coords(r)
        s0 = orc.block(n[0], w.np_[0], c[0])[0]
        s1 = orc.block(n[1], w.np_[1], c[1])[0]
        blocks.append(outs[r][:s0 * s1 * n[2]].reshape(s0, s1, n[2]).copy())
    return blocks


SYMBOL_OPS = {"inv_laplace": 5, "inv_biharmonic": 6, "lin": 7, "inv": 8}


def symbol(w, x_global, op, a=0.0, b=0.0):
    ins = _local_real(w, x_global)
    outs = w.alloc_out()
    rc = lib().emu_execute_symbol(w.h, SYMBOL_OPS[op], a, b, _ptrs(ins), _ptrs(outs))
    if rc:
        raise RuntimeError(f"emu_execute_symbol rc={rc}")
    return _blocks_real(w, outs)


def axis_r2c(w, x_global, axis):
    """single-axis forward: per-rank half spectra in the reference's osize_0 / osize_yi / osize_xi layouts"""
    ins = _local_real(w, x_global)
    outs = w.alloc_out()
    rc = lib().emu_execute_axis(w.h, 0, axis, _ptrs(ins), _ptrs(outs))
    if rc:
        raise RuntimeError(f"emu_execute_axis rc={rc}")
    Xa = orc.axis_forward(x_global, axis)
    res = []
    for r in range(w.nprocs):
        ref = orc.axis_local_spectrum(Xa, w.n, w.np_, w.coords(r), axis)
        res.append((outs[r].view(w.cdt)[:ref.size].reshape(ref.shape).copy(), ref))
    return res


def axis_c2r(w, x_global, axis):
    """single-axis inverse of the oracle's local spectra -> real blocks"""
    Xa = orc.axis_forward(x_global, axis)
    ins = []
    for r in range(w.nprocs):
        loc = orc.axis_local_spectrum(Xa, w.n, w.np_, w.coords(r), axis).astype(w.cdt)
        ins.append(np.ascontiguousarray(loc) if loc.size else np.zeros(2, dtype=w.cdt))
    outs = w.alloc_out()
    rc = lib().emu_execute_axis(w.h, 1, axis, _ptrs(ins), _ptrs(outs))
    if rc:
        raise RuntimeError(f"emu_execute_axis rc={rc}")
    return _blocks_real(w, outs)


def axis_deriv(w, x_global, axis, accumulate_onto=None):
    ins = _local_real(w, x_global)
    outs = w.alloc_out()
    if accumulate_onto is not None:
        for r, b in enumerate(_local_real(w, accumulate_onto)):
            outs[r][:b.size] = b.ravel()
    rc = lib().emu_execute_axis(w.h, 3 if accumulate_onto is not None else 2, axis, _ptrs(ins), _ptrs(outs))
    if rc:
        raise RuntimeError(f"emu_execute_axis rc={rc}")
    return _blocks_real(w, outs)


def grad(w, x_global, comps=7, fast=False):
    ins = _local_real(w, x_global)
    outs = [w.alloc_out() for _ in range(3)]
    fn = lib().emu_execute_grad_axes if fast else lib().emu_execute_grad
    rc = fn(w.h, _ptrs(ins), _ptrs(outs[0]), _ptrs(outs[1]), _ptrs(outs[2]), comps)
    if rc:
        raise RuntimeError(f"emu_execute_grad rc={rc}")
    return [_blocks_real(w, o) for o in outs]


def divergence(w, ax, ay, az, fast=False):
    ins = [_local_real(w, v) for v in (ax, ay, az)]
    out = w.alloc_out()
    fn = lib().emu_execute_div_axes if fast else lib().emu_execute_div
    rc = fn(w.h, _ptrs(ins[0]), _ptrs(ins[1]), _ptrs(ins[2]), _ptrs(out))
    if rc:
        raise RuntimeError(f"emu_execute_div rc={rc}")
    return _blocks_real(w, out)


if __name__ == "__main__":
    cases = [((8, 8, 8), (1, 1)), ((16, 8, 12), (1, 1)), ((8, 8, 8), (1, 2)), ((8, 8, 8), (2, 2)),
             ((12, 10, 16), (2, 4)), ((6, 12, 10), (3, 2)), ((5, 6, 6), (4, 2)), ((16, 12, 6), (2, 4)),
             ((3, 5, 4), (4, 4)), ((32, 24, 20), (2, 2)), ((64, 48, 40), (1, 2))]
    for n, np_ in cases:
        x = orc.synthetic_real(n, 7)
        w = World(n, np_, r2c=True)
        X = orc.forward_r2c(x)
        e1 = orc.rel_l2(orc.assemble_output(w.forward(x), n, np_), X)
        e1p = orc.rel_l2(orc.assemble_output(w.forward(x, padded=True), n, np_), X)
        xb = orc.assemble_input(w.inverse(X), n, np_)
        e2 = orc.rel_l2(xb, x * x.size)
        lap = orc.rel_l2(orc.assemble_input(w.spectral(x, "laplace"), n, np_), orc.laplace(x))
        w.close()
        z = orc.synthetic_complex(n, 8)
        w = World(n, np_, r2c=False)
        Z = orc.forward_c2c(z)
        e3 = orc.rel_l2(orc.assemble_output(w.forward(z), n, np_, r2c=False), Z)
        zb = orc.assemble_input(w.inverse(Z), n, np_)
        e4 = orc.rel_l2(zb, z * z.size)
        w.close()
        print(n, np_, "r2c %.2e (padded %.2e) c2r %.2e lap %.2e | c2c fwd %.2e inv %.2e" % (e1, e1p, e2, lap, e3, e4))
