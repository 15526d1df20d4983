"""ctypes host layer over libaccfft_b200.so, mirroring AccFFT's GPU API names and argument meaning.

Reference interfaces mirrored (paths in the AccFFT tree):
  accfft_create_comm                         src/accfft_common.cpp:275-305
  accfft_local_size_dft_{r2c,c2c}_gpu[f]     src/accfft_gpu.cpp:62-65, :689-692
  accfft_plan_dft_3d_{r2c,c2c}_gpu[f]        src/accfft_gpu.cpp:77-442, :704-931
  accfft_execute_{r2c,c2r,c2c}_gpu[f]        src/accfft_gpu.cpp:635-675, :943-1136
  accfft_{grad,laplace,divergence,biharmonic}_gpu[f]   src/operators_gpu.cpp:85-135
  accfft_destroy_plan_gpu                    src/accfft_gpu.cpp:1150-1192

Data arguments are torch CUDA tensors (or raw integer device pointers); torch is used for device
memory, streams and -- for multi-rank plans -- torch.distributed as the transport that all-gathers
the CUDA-IPC blobs at plan time.  All arithmetic happens in the CUDA library; nothing here computes.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

ACCFFT_FORWARD = -1
ACCFFT_BACKWARD = +1
ACCFFT_ESTIMATE = 1
ACCFFT_MEASURE = 2
ACCFFT_PATIENT = 4

_R2C, _C2C = 0, 1
_DOUBLE, _FLOAT = 0, 1
_BLOB = 128

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class AccfftError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get("ACCFFT_B200_LIB", os.path.join(_HERE, "libaccfft_b200.so"))


def lib():
    """Load libaccfft_b200.so (built in-tree by accfft_b200/csrc/Makefile).  Fails loudly if missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise AccfftError(
            f"{path} not found: build it with `make -C accfft_b200/csrc -j8` (or __graft_entry__.build()). "
            "accfft_b200 has no CPU fallback.")
    L = ctypes.CDLL(path)
    c_int, c_ll, c_vp, c_dp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(c_int)
    L.accfftb_last_error.restype = ctypes.c_char_p
    L.accfftb_version.restype = ctypes.c_char_p
    L.accfftb_create_comm_dims.argtypes = [c_int, ip]
    L.accfftb_rank_to_coords.argtypes = [c_int, ip, ip]
    L.accfftb_rank_to_coords.restype = None
    for f in (L.accfftb_local_size_dft_r2c, L.accfftb_local_size_dft_c2c):
        f.restype = c_ll
        f.argtypes = [ip, ip, ip, c_int, ip, ip, ip, ip]
    L.accfftb_stage_size.argtypes = [ip, ip, ip, c_int, c_int, c_int, ip, ip]
    L.accfftb_length_supported.argtypes = [c_int]
    L.accfftb_length_fast.argtypes = [c_int]
    L.accfftb_plan_create.restype = c_vp
    L.accfftb_plan_create.argtypes = [ip, ip, c_int, c_int, c_int, ctypes.c_uint]
    L.accfftb_plan_ipc_export.argtypes = [c_vp, c_vp]
    L.accfftb_plan_ipc_import.argtypes = [c_vp, c_vp, c_int]
    L.accfftb_plan_attach_local_peers.argtypes = [c_vp, ctypes.POINTER(c_vp), c_int]
    L.accfftb_plan_set_stream.argtypes = [c_vp, c_vp]
    L.accfftb_plan_set_blocking.argtypes = [c_vp, c_int]
    L.accfftb_plan_set_profiling.argtypes = [c_vp, c_int]
    L.accfftb_plan_profile_read.argtypes = [c_vp, c_dp, ctypes.POINTER(c_ll)]
    L.accfftb_plan_stage_variant.argtypes = [c_vp, c_int]
    L.accfftb_plan_fused_info.argtypes = [c_vp, c_int, ip]
    L.accfftb_plan_create_many.restype = c_vp
    L.accfftb_plan_create_many.argtypes = [ip, ip, c_int, c_int, c_int, ctypes.c_uint, c_int]
    L.accfftb_plan_howmany.argtypes = [c_vp]
    vpp = ctypes.POINTER(c_vp)
    L.accfftb_execute_r2c_many.argtypes = [c_vp, c_int, vpp, vpp, c_dp]
    L.accfftb_execute_c2r_many.argtypes = [c_vp, c_int, vpp, vpp, c_dp]
    L.accfftb_execute_c2c_many.argtypes = [c_vp, c_int, c_int, vpp, vpp, c_dp]
    L.accfftb_plan_alloc_max.restype = c_ll
    L.accfftb_plan_alloc_max.argtypes = [c_vp]
    L.accfftb_plan_launch_count.restype = c_ll
    L.accfftb_plan_launch_count.argtypes = [c_vp]
    L.accfftb_plan_copy_count.restype = c_ll
    L.accfftb_plan_copy_count.argtypes = [c_vp]
    L.accfftb_plan_exchange_mode.argtypes = [c_vp, ip, ip]
    L.accfftb_plan_check.argtypes = [c_vp]
    L.accfftb_plan_destroy.argtypes = [c_vp]
    L.accfftb_execute_r2c.argtypes = [c_vp, c_vp, c_vp, c_dp, c_int]
    L.accfftb_execute_c2r.argtypes = [c_vp, c_vp, c_vp, c_dp, c_int]
    L.accfftb_execute_c2c.argtypes = [c_vp, c_int, c_vp, c_vp, c_dp, c_int]
    L.accfftb_laplace.argtypes = [c_vp, c_vp, c_vp, c_dp]
    L.accfftb_biharmonic.argtypes = [c_vp, c_vp, c_vp, c_dp]
    L.accfftb_grad.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_dp]
    L.accfftb_divergence.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_dp]
    L.accfftb_inv_laplace.argtypes = [c_vp, c_vp, c_vp, c_dp]
    L.accfftb_inv_biharmonic.argtypes = [c_vp, c_vp, c_vp, c_dp]
    L.accfftb_symbol_solve.argtypes = [c_vp, c_vp, c_vp, ctypes.c_double, ctypes.c_double, c_int, c_dp]
    L.accfftb_grad_slow.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_dp]
    L.accfftb_divergence_slow.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_dp]
    L.accfftb_execute_r2c_axis.argtypes = [c_vp, c_int, c_vp, c_vp, c_dp]
    L.accfftb_execute_c2r_axis.argtypes = [c_vp, c_int, c_vp, c_vp, c_dp]
    _LIB = L
    return L


def _err() -> str:
    return lib().accfftb_last_error().decode("utf-8", "replace")


def _check(rc: int, what: str):
    if rc != 0:
        raise AccfftError(f"{what} failed (rc={rc}): {_err()}")


def _i3(v: Sequence[int]):
    return (ctypes.c_int * 3)(*[int(x) for x in v])


def _i2(v: Sequence[int]):
    return (ctypes.c_int * 2)(*[int(x) for x in v])


def length_supported(n: int) -> bool:
    """can an axis of this length be planned (scheduled kernel or the generic O(n^2) fallback)?"""
    return bool(lib().accfftb_length_supported(int(n)))


def length_fast(n: int) -> bool:
    """does a compiled Stockham schedule exist for this complex length?"""
    return bool(lib().accfftb_length_fast(int(n)))


class Comm:
    """Stand-in for the Cartesian communicator accfft_create_comm returns (MPI is not used here).

    rank / nprocs / c_dims describe the P0 x P1 grid; ``group`` is the torch.distributed process group
    (None = default group) used ONLY as a transport for the plan-time exchange of CUDA IPC handles.
    """

    def __init__(self, rank: int, nprocs: int, c_dims: Sequence[int], group=None):
        self.rank, self.nprocs, self.c_dims, self.group = int(rank), int(nprocs), (int(c_dims[0]), int(c_dims[1])), group
        if self.c_dims[0] * self.c_dims[1] != self.nprocs:
            raise AccfftError("c_dims does not match the number of ranks")
        c = _i2((0, 0))
        lib().accfftb_rank_to_coords(self.rank, _i2(self.c_dims), c)
        self.coords = (c[0], c[1])

    def all_gather_bytes(self, payload: bytes) -> List[bytes]:
        if self.nprocs == 1:
            return [payload]
        import torch
        import torch.distributed as dist
        backend = dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
        out = [torch.empty_like(mine) for _ in range(self.nprocs)]
        dist.all_gather(out, mine, group=self.group)
        return [bytes(t.cpu().numpy().tobytes()) for t in out]

    def barrier(self):
        if self.nprocs > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)


def accfft_create_comm(c_dims: List[int], group=None, rank: Optional[int] = None, nprocs: Optional[int] = None) -> Comm:
    """accfft_create_comm(in_comm, c_dims, &c_comm): like the reference, OVERWRITES c_dims in place when
    c_dims[0]*c_dims[1] != nprocs (MPI_Dims_create fallback, src/accfft_common.cpp:281-289)."""
    if rank is None or nprocs is None:
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, nprocs = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, nprocs = 0, 1
        except ImportError:
            rank, nprocs = 0, 1
    d = _i2(c_dims)
    if lib().accfftb_create_comm_dims(int(nprocs), d):
        c_dims[0], c_dims[1] = d[0], d[1]
    return Comm(rank, nprocs, (d[0], d[1]), group)


def _local_size(fn, n, comm: Comm, sizeof_real: int):
    isz, ist, osz, ost = _i3((0,) * 3), _i3((0,) * 3), _i3((0,) * 3), _i3((0,) * 3)
    alloc = fn(_i3(n), _i2(comm.c_dims), _i2(comm.coords), sizeof_real, isz, ist, osz, ost)
    if alloc < 0:
        return -1, list(isz), list(ist), list(osz), list(ost)   # the reference also returns -1 (overflow guard)
    return int(alloc), list(isz), list(ist), list(osz), list(ost)


def accfft_local_size_dft_r2c_gpu(n, comm: Comm):
    """-> (alloc_max bytes, isize, istart, osize, ostart)"""
    return _local_size(lib().accfftb_local_size_dft_r2c, n, comm, 8)


def accfft_local_size_dft_r2c_gpuf(n, comm: Comm):
    return _local_size(lib().accfftb_local_size_dft_r2c, n, comm, 4)


def accfft_local_size_dft_c2c_gpu(n, comm: Comm):
    return _local_size(lib().accfftb_local_size_dft_c2c, n, comm, 8)


def accfft_local_size_dft_c2c_gpuf(n, comm: Comm):
    return _local_size(lib().accfftb_local_size_dft_c2c, n, comm, 4)


def _ptr(t) -> int:
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    if not t.is_cuda:
        raise AccfftError("accfft_b200 takes device (CUDA) tensors; there is no CPU path")
    if not t.is_contiguous():
        raise AccfftError("tensor must be contiguous")
    return t.data_ptr()


class Plan:
    """accfft_plan_gpu / accfft_plan_gpuf.  flags defaults to ACCFFT_MEASURE like the reference
    (include/accfft_gpu.h:103): plan creation times the two strided-stage kernels on plan-owned scratch and keeps the
    faster one per stage; ACCFFT_ESTIMATE skips the timing and uses a size heuristic."""

    def __init__(self, n, comm: Comm, r2c: bool, single: bool, flags: int = ACCFFT_MEASURE, wire: bool = True, howmany: int = 1):
        """howmany > 1: a batched plan (accfftb_plan_create_many) -- howmany sets of exchange buffers, execute_*_many take up
        to howmany fields per call and share the inter-GPU barriers between them"""
        self.n, self.comm, self.r2c, self.single = tuple(int(x) for x in n), comm, r2c, single
        self.howmany = int(howmany)
        self._local_group = None
        L = lib()
        self._h = L.accfftb_plan_create_many(_i3(n), _i2(comm.c_dims), comm.rank, _R2C if r2c else _C2C,
                                             _FLOAT if single else _DOUBLE, flags, self.howmany)
        err = None if self._h else "plan creation failed: " + _err()
        # plan creation is collective: agree on the outcome before anybody waits for anybody's IPC blob (the C++ layer does
        # the same with an MPI_Allreduce(MIN), csrc/accfft_gpu_api.cpp), so that a failure on one rank raises on ALL ranks
        # instead of leaving the others blocked in the all-gather below
        collective = comm.nprocs > 1 and wire
        if collective:
            oks = comm.all_gather_bytes(b"\x01" if err is None else b"\x00")
            if err is None and any(o != b"\x01" for o in oks):
                err = "plan creation failed on rank(s) %s" % [i for i, o in enumerate(oks) if o != b"\x01"]
        if err is not None:
            if self._h:
                L.accfftb_plan_destroy(self._h)
                self._h = None
            # the reference prints and returns NULL (src/accfft_gpu.cpp:187-191); raise with the message instead
            raise AccfftError(err)
        sz = (accfft_local_size_dft_r2c_gpuf if single else accfft_local_size_dft_r2c_gpu) if r2c else \
             (accfft_local_size_dft_c2c_gpuf if single else accfft_local_size_dft_c2c_gpu)
        self.alloc_max, self.isize, self.istart, self.osize, self.ostart = sz(n, comm)
        if collective:
            blob = ctypes.create_string_buffer(_BLOB)
            rc = L.accfftb_plan_ipc_export(self._h, blob)
            blobs = b"".join(comm.all_gather_bytes(blob.raw))
            if rc == 0:
                buf = ctypes.create_string_buffer(blobs, len(blobs))
                rc = L.accfftb_plan_ipc_import(self._h, buf, comm.nprocs)
            msg = _err() if rc else ""
            oks = comm.all_gather_bytes(b"\x01" if rc == 0 else b"\x00")
            if any(o != b"\x01" for o in oks):
                L.accfftb_plan_destroy(self._h)
                self._h = None
                raise AccfftError("wiring the exchange buffers failed on rank(s) %s%s" %
                                  ([i for i, o in enumerate(oks) if o != b"\x01"], (": " + msg) if msg else ""))
            comm.barrier()

    # ---- configuration -------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        _check(lib().accfftb_plan_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "set_stream")

    def set_blocking(self, blocking: bool):
        _check(lib().accfftb_plan_set_blocking(self._h, int(blocking)), "set_blocking")

    def set_profiling(self, on: bool):
        _check(lib().accfftb_plan_set_profiling(self._h, int(on)), "set_profiling")

    PROFILE_SLOTS = ("fwd_z", "fwd_y", "fwd_x", "inv_x", "inv_y", "inv_z", "barrier", "axis_copy", "axis_fft", "ce_copy")

    def profile_read(self):
        """-> {slot name: (total ms, launches)} since the last read (synchronises the plan's stream)"""
        ms = (ctypes.c_double * len(self.PROFILE_SLOTS))()
        cnt = (ctypes.c_longlong * len(self.PROFILE_SLOTS))()
        _check(lib().accfftb_plan_profile_read(self._h, ms, cnt), "profile_read")
        return {name: (ms[i], int(cnt[i])) for i, name in enumerate(self.PROFILE_SLOTS)}

    def stage_kernels(self):
        """-> {stage: "plain" | "tma" | "heuristic"} for the four strided stages (see accfftb_plan_stage_variant)"""
        names = {0: "heuristic", 1: "plain", 2: "tma", 3: "wide"}
        return {self.PROFILE_SLOTS[s]: names.get(lib().accfftb_plan_stage_variant(self._h, s), "?") for s in (1, 2, 3, 4)}

    def fused_info(self, inverse: bool = False):
        """accfftb_plan_fused_info: None if the last transform in that direction ran its local Z / Y passes as two kernels, else
        what the fused persistent kernel used: {"cp", "lag", "ns", "chunks", "ring_planes", "grid", "threads", "smem"}"""
        arr = (ctypes.c_int * 8)()
        r = lib().accfftb_plan_fused_info(self._h, 1 if inverse else 0, arr)
        if r != 1:
            return None
        return dict(zip(("cp", "lag", "ns", "chunks", "ring_planes", "grid", "threads", "smem"), [int(v) for v in arr]))

    @property
    def launch_count(self) -> int:
        return int(lib().accfftb_plan_launch_count(self._h))

    @property
    def copy_count(self) -> int:
        """copy-engine transfers enqueued so far (slab-grid exchange)"""
        return int(lib().accfftb_plan_copy_count(self._h))

    def exchange_mode(self):
        """-> {"mode": "copy_engine" | "peer_store" | "local", "chunks": n, "lanes": n}"""
        ch, ln = ctypes.c_int(0), ctypes.c_int(0)
        m = lib().accfftb_plan_exchange_mode(self._h, ctypes.byref(ch), ctypes.byref(ln))
        mode = "copy_engine" if m == 1 else ("peer_store" if self.comm.nprocs > 1 else "local")
        return {"mode": mode, "chunks": ch.value, "lanes": ln.value}

    def check(self):
        """accfftb_plan_check: raises once an inter-GPU barrier of the plan has timed out (sticky); never synchronises"""
        _check(lib().accfftb_plan_check(self._h), "accfftb_plan_check")

    # ---- transforms ----------------------------------------------------------------------------
    @staticmethod
    def _timer(timer):
        if timer is None:
            return None, None
        arr = (ctypes.c_double * max(7, len(timer)))(*timer)
        return arr, ctypes.cast(arr, ctypes.POINTER(ctypes.c_double))

    @staticmethod
    def _timer_back(arr, timer):
        if timer is not None:
            for i in range(len(timer)):
                timer[i] = arr[i]

    def execute_r2c(self, data, data_out, timer=None, xyz: int = 0b111):
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_r2c(self._h, _ptr(data), _ptr(data_out), p, xyz), "accfft_execute_r2c_gpu")
        self._timer_back(arr, timer)

    def execute_c2r(self, data, data_out, timer=None, xyz: int = 0b111):
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_c2r(self._h, _ptr(data), _ptr(data_out), p, xyz), "accfft_execute_c2r_gpu")
        self._timer_back(arr, timer)

    def execute_c2c(self, direction: int, data, data_out, timer=None, xyz: int = 0b111):
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_c2c(self._h, direction, _ptr(data), _ptr(data_out), p, xyz),
               "accfft_execute_c2c_gpu")
        self._timer_back(arr, timer)

    # ---- single-axis real transforms (accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t, src/accfft_gpu.cpp:1212-1604) -------
    AXES = {"x": 0, "y": 1, "z": 2}

    # ---- batched fields (accfftb_execute_*_many) -------------------------------------------------
    @staticmethod
    def _ptr_array(ts):
        arr = (ctypes.c_void_p * len(ts))()
        for i, t in enumerate(ts):
            arr[i] = _ptr(t)
        return arr

    def execute_r2c_many(self, datas, data_outs, timer=None):
        """howmany = len(datas) real fields -> their spectra, through one set of exchanges"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_r2c_many(self._h, len(datas), self._ptr_array(datas), self._ptr_array(data_outs), p),
               "accfftb_execute_r2c_many")
        self._timer_back(arr, timer)

    def execute_c2r_many(self, datas, data_outs, timer=None):
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_c2r_many(self._h, len(datas), self._ptr_array(datas), self._ptr_array(data_outs), p),
               "accfftb_execute_c2r_many")
        self._timer_back(arr, timer)

    def execute_c2c_many(self, direction: int, datas, data_outs, timer=None):
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_c2c_many(self._h, direction, len(datas), self._ptr_array(datas), self._ptr_array(data_outs), p),
               "accfftb_execute_c2c_many")
        self._timer_back(arr, timer)

    def axis_size(self, axis):
        """-> (size, start) of this rank's block of the single-axis half spectrum (osize_xi / osize_yi / osize_0)"""
        axis = self.AXES.get(axis, axis)
        which = {0: 6, 1: 4, 2: 0}[axis]
        size, start = (ctypes.c_int * 3)(), (ctypes.c_int * 3)()
        _check(lib().accfftb_stage_size(_i3(self.n), _i2(self.comm.c_dims), _i2(self.comm.coords), _R2C,
                                        4 if self.single else 8, which, size, start), "accfftb_stage_size")
        if axis == 1:   # starts of the y path: ostart_y (which = 3)
            s3 = (ctypes.c_int * 3)()
            _check(lib().accfftb_stage_size(_i3(self.n), _i2(self.comm.c_dims), _i2(self.comm.coords), _R2C,
                                            4 if self.single else 8, 3, s3, start), "accfftb_stage_size")
        return tuple(size), tuple(start)

    def execute_r2c_axis(self, axis, data, data_out, timer=None):
        """accfft_execute_r2c_{x,y,z}_gpu_t(plan, data, data_out, timer): axis = 0/1/2 or "x"/"y"/"z" """
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_r2c_axis(self._h, self.AXES.get(axis, axis), _ptr(data), _ptr(data_out), p),
               "accfft_execute_r2c_axis")
        self._timer_back(arr, timer)

    def execute_c2r_axis(self, axis, data, data_out, timer=None):
        """accfft_execute_c2r_{x,y,z}_gpu_t(plan, data, data_out, timer)"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_execute_c2r_axis(self._h, self.AXES.get(axis, axis), _ptr(data), _ptr(data_out), p),
               "accfft_execute_c2r_axis")
        self._timer_back(arr, timer)

    # ---- spectral operators (accfft_operators_gpu.h) ---------------------------------------------
    def laplace(self, LA, A, timer=None):
        """accfft_laplace_gpu(LA, A, plan, timer)"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_laplace(self._h, _ptr(A), _ptr(LA), p), "accfft_laplace_gpu")
        self._timer_back(arr, timer)

    def biharmonic(self, BA, A, timer=None):
        """accfft_biharmonic_gpu(BA, A, plan, timer)"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_biharmonic(self._h, _ptr(A), _ptr(BA), p), "accfft_biharmonic_gpu")
        self._timer_back(arr, timer)

    def grad(self, Ax, Ay, Az, A, xyz: int = 0b111, timer=None):
        """accfft_grad_gpu(Ax, Ay, Az, A, plan, &XYZ, timer); xyz bit 0 = x, bit 1 = y, bit 2 = z"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_grad(self._h, _ptr(A), _ptr(Ax), _ptr(Ay), _ptr(Az), xyz, p), "accfft_grad_gpu")
        self._timer_back(arr, timer)

    def divergence(self, div, Ax, Ay, Az, timer=None):
        """accfft_divergence_gpu(div, Ax, Ay, Az, plan, timer)"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_divergence(self._h, _ptr(Ax), _ptr(Ay), _ptr(Az), _ptr(div), p), "accfft_divergence_gpu")
        self._timer_back(arr, timer)

    def inv_laplace(self, invLA, A, timer=None):
        """GPU twin of accfft_inv_laplace(invLA, A, plan, timer) (CPU-only in the reference, src/operators.txx:275-333)"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_inv_laplace(self._h, _ptr(A), _ptr(invLA), p), "accfft_inv_laplace_gpu")
        self._timer_back(arr, timer)

    def inv_biharmonic(self, invBA, A, timer=None):
        """GPU twin of accfft_inv_biharmonic (src/operators.txx:335-392)"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_inv_biharmonic(self._h, _ptr(A), _ptr(invBA), p), "accfft_inv_biharmonic_gpu")
        self._timer_back(arr, timer)

    def symbol_solve(self, out, A, alpha: float, beta: float, inverse: bool, timer=None):
        """out = F^-1[(alpha + beta |k|^2)^(+-1) F[A]]: Poisson filter (steps/step4), heat steps (steps/step6), Helmholtz"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_symbol_solve(self._h, _ptr(A), _ptr(out), float(alpha), float(beta), int(bool(inverse)), p),
               "accfftb_symbol_solve")
        self._timer_back(arr, timer)

    def grad_slow(self, Ax, Ay, Az, A, xyz: int = 0b111, timer=None):
        """accfft_grad_gpu_slow_t (src/operators_gpu.txx:72-170): full 3-D transforms instead of per-axis ones"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_grad_slow(self._h, _ptr(A), _ptr(Ax), _ptr(Ay), _ptr(Az), xyz, p), "accfft_grad_slow_gpu")
        self._timer_back(arr, timer)

    def divergence_slow(self, div, Ax, Ay, Az, timer=None):
        """accfft_divergence_slow_gpu (src/operators_gpu.txx:336-446)"""
        arr, p = self._timer(timer)
        _check(lib().accfftb_divergence_slow(self._h, _ptr(Ax), _ptr(Ay), _ptr(Az), _ptr(div), p),
               "accfft_divergence_slow_gpu")
        self._timer_back(arr, timer)

    # ---- teardown --------------------------------------------------------------------------------
    def destroy(self):
        """accfft_destroy_plan_gpu(plan)"""
        if getattr(self, "_h", None):
            if self.comm.nprocs > 1 and self._local_group is None:
                try:
                    import torch
                    torch.cuda.synchronize()
                    self.comm.barrier()   # nobody unmaps a slab a peer may still be writing
                except Exception:
                    pass
            lib().accfftb_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                if self.comm.nprocs == 1:
                    lib().accfftb_plan_destroy(self._h)
                    self._h = None
                else:
                    # destroying a multi-rank plan is collective (peers may still write into this rank's buffers): it cannot
                    # be done from a finaliser.  Say so instead of leaking three exchange buffers silently.
                    import warnings
                    warnings.warn("accfft_b200.Plan of a %d-rank communicator was garbage collected without destroy(): its "
                                  "exchange buffers and IPC mappings are leaked" % self.comm.nprocs, ResourceWarning)
        except Exception:
            pass


def attach_local_peers(plans: Sequence[Plan]):
    """Wire plans whose ranks all live in THIS process (one host thread per rank, created with wire=False):
    peers' exchange slabs are addressed directly (cudaDeviceEnablePeerAccess between devices) instead of
    through CUDA IPC.  plans must be in rank order."""
    arr = (ctypes.c_void_p * len(plans))(*[p._h for p in plans])
    for p in plans:
        _check(lib().accfftb_plan_attach_local_peers(p._h, arr, len(plans)), "accfftb_plan_attach_local_peers")
        p._local_group = list(plans)


def accfft_plan_dft_3d_r2c_gpu(n, comm: Comm, flags: int = ACCFFT_MEASURE) -> Plan:
    return Plan(n, comm, r2c=True, single=False, flags=flags)


def accfft_plan_dft_3d_r2c_gpuf(n, comm: Comm, flags: int = ACCFFT_MEASURE) -> Plan:
    return Plan(n, comm, r2c=True, single=True, flags=flags)


def accfft_plan_dft_3d_c2c_gpu(n, comm: Comm, flags: int = ACCFFT_MEASURE) -> Plan:
    return Plan(n, comm, r2c=False, single=False, flags=flags)


def accfft_plan_dft_3d_c2c_gpuf(n, comm: Comm, flags: int = ACCFFT_MEASURE) -> Plan:
    return Plan(n, comm, r2c=False, single=True, flags=flags)
