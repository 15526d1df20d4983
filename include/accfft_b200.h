/* accfft_b200.h -- C ABI of libaccfft_b200.so: the B200-native engine behind AccFFT's GPU API.
 *
 * Plain C, plain pointers and sizes.  Every entry point names the reference interface it replaces
 * (paths relative to the AccFFT source tree).  The C++ header include/accfft_gpu.h re-creates the
 * reference's own C++ API (same names / argument meaning) on top of these calls; accfft_b200/api.py
 * is the ctypes binding used by the tests and the benchmark.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers on the CUDA device that was current when the plan was
 *     created (one rank <-> one GPU, chosen by the caller exactly as with the reference, which never
 *     calls cudaSetDevice);
 *   - calls on a plan are collective over the plan's P0 x P1 process grid;
 *   - return value 0 = success; otherwise a negative accfftb error or a positive cudaError_t, with
 *     accfftb_last_error() returning a message for the calling thread;
 *   - transforms are unnormalised (c2r(r2c(x)) = N0*N1*N2*x) like the reference;
 *   - prec: 0 = double (accfft_*_gpu), 1 = float (accfft_*_gpuf).
 */
#ifndef ACCFFT_B200_H
#define ACCFFT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct accfftb_plan accfftb_plan;

#define ACCFFTB_FORWARD (-1) /* ACCFFT_FORWARD,  include/accfft_common.h:38 */
#define ACCFFTB_BACKWARD (+1) /* ACCFFT_BACKWARD, include/accfft_common.h:39 */
#define ACCFFTB_ESTIMATE 1    /* ACCFFT_ESTIMATE, include/accfft_common.h:40 */
#define ACCFFTB_MEASURE 2     /* ACCFFT_MEASURE,  include/accfft_common.h:41 */
#define ACCFFTB_PATIENT 4     /* ACCFFT_PATIENT,  include/accfft_common.h:42 */

#define ACCFFTB_R2C 0
#define ACCFFTB_C2C 1
#define ACCFFTB_DOUBLE 0
#define ACCFFTB_FLOAT 1

#define ACCFFTB_IPC_BLOB_BYTES 128 /* per-rank blob exchanged by accfftb_plan_ipc_export / _import */
#define ACCFFTB_MAX_GROUP 16       /* max ranks along one dimension of the process grid */

#define ACCFFTB_ERR_ARG (-1)         /* bad argument */
#define ACCFFTB_ERR_UNSUPPORTED (-2) /* no kernel schedule for a transform length, odd N2 for r2c, ... */
#define ACCFFTB_ERR_STATE (-3)       /* e.g. multi-rank plan executed before accfftb_plan_ipc_import */
#define ACCFFTB_ERR_TIMEOUT (-4)     /* inter-GPU barrier timed out (a peer died or diverged) */
#define ACCFFTB_ERR_OVERFLOW (-5)    /* local block exceeds 2^31 elements: src/accfft_common.cpp:179-184 */

const char* accfftb_last_error(void);
const char* accfftb_version(void);

/* ---- process grid and pencil decomposition (host integer math only; no GPU needed) ------------ */

/* The MPI_Dims_create fallback of accfft_create_comm (src/accfft_common.cpp:281-289): if
 * c_dims[0]*c_dims[1] != nprocs the dims are OVERWRITTEN with a balanced, non-increasing
 * factorisation (2 -> 2x1, 4 -> 2x2, 8 -> 4x2).  Returns 1 if c_dims was overwritten, 0 if kept. */
int accfftb_create_comm_dims(int nprocs, int c_dims[2]);

/* Cartesian rank -> (coord0, coord1), row-major as MPI_Cart_create(reorder ignored) lays them out
 * (src/accfft_common.cpp:298): rank = coord0 * c_dims[1] + coord1. */
void accfftb_rank_to_coords(int rank, const int c_dims[2], int coords[2]);

/* accfft_local_size_dft_r2c_gpu / _gpuf  (src/accfft_gpu.cpp:62-65 -> src/accfft_common.cpp:98-192).
 * sizeof_real = 8 (double API) or 4 (float API).  Returns alloc_max in BYTES (64-bit here; the
 * reference returns int) or ACCFFTB_ERR_OVERFLOW. */
long long accfftb_local_size_dft_r2c(const int n[3], const int c_dims[2], const int coords[2], int sizeof_real,
                                     int isize[3], int istart[3], int osize[3], int ostart[3]);

/* accfft_local_size_dft_c2c_gpu / _gpuf  (src/accfft_gpu.cpp:689-692 -> src/accfft_common.cpp:203-256). */
long long accfftb_local_size_dft_c2c(const int n[3], const int c_dims[2], const int coords[2], int sizeof_real,
                                     int isize[3], int istart[3], int osize[3], int ostart[3]);

/* Stage geometry the reference keeps in the public plan struct (include/accfft_gpu.h:37-95):
 * which = 0,1,2 -> osize_0/1/2 + ostart_0/1/2;  3 -> osize_y/ostart_y;  4 -> osize_yi;  5 -> osize_x;
 * 6 -> osize_xi.  start may be NULL. */
int accfftb_stage_size(const int n[3], const int c_dims[2], const int coords[2], int kind, int sizeof_real,
                       int which, int size[3], int start[3]);

/* Transform lengths.  Like the reference (which hands any length to cuFFT, src/accfft_gpu.cpp:173-253) every length plans:
 *   accfftb_length_fast(n)       1 if a compiled Stockham schedule exists for the complex length n: 2^a 3^b 5^c <= 2048
 *                                (b, c <= 3); 7 * 2^a 3^b 5^c (b, c <= 1); 11 / 13 / 49 times a power of two; 4096.  The
 *                                r2c / c2r Z pass runs the N2/2-point schedule: it needs N2 even with N2/2 fast.
 *   accfftb_length_supported(n)  1 if an axis of length n can be planned at all: fast, or through the generic O(n^2) stage
 *                                kernel (any n up to 7168; also serves odd N2 in r2c plans).  That kernel is a correctness
 *                                fallback for the odd axis of an awkward grid, not a fast path.
 * The single-axis executes (accfftb_execute_{r2c,c2r}_axis) exist for fast lengths only; the per-axis operators fall back to
 * their full 3-D scheme (same operator) when an axis is not fast. */
int accfftb_length_supported(int n);
int accfftb_length_fast(int n);

/* ---- plans ------------------------------------------------------------------------------------ */

/* accfft_plan_dft_3d_r2c_gpu / _gpuf (src/accfft_gpu.cpp:77-442) and accfft_plan_dft_3d_c2c_gpu / _gpuf
 * (:704-931).  kind = ACCFFTB_R2C / ACCFFTB_C2C.  The plan owns two exchange buffers of alloc_max bytes
 * on the current device (the reference's Mem_Mgr_gpu owns three device + two pinned host buffers,
 * src/transpose_gpu.cpp:58-131) and the twiddle tables.  Unlike the reference, plan creation never
 * touches user data (ACCFFT_MEASURE clobbers data_out there, src/accfft_gpu.cpp:382-386) so no data
 * pointers are taken.  flags: ACCFFTB_ESTIMATE or ACCFFTB_MEASURE / ACCFFTB_PATIENT (see accfftb_plan_stage_variant).
 * Returns NULL on failure (as the reference does, :187-191). */
accfftb_plan* accfftb_plan_create(const int n[3], const int c_dims[2], int rank, int kind, int prec, unsigned flags);

/* Multi-GPU wiring (replaces MPI_Comm_split + the host-staged MPI_Alltoall[v] machinery of
 * src/transpose_gpu.cpp:3940-3955): each rank exports a blob (CUDA IPC handle of its exchange slab),
 * the caller all-gathers the blobs over ANY transport (torch.distributed, MPI, files ...) and hands the
 * rank-ordered array to _import, which maps every peer's slab over NVLink.  Single-rank plans skip both. */
int accfftb_plan_ipc_export(accfftb_plan* plan, void* blob);
int accfftb_plan_ipc_import(accfftb_plan* plan, const void* blobs, int nblobs);
/* Same wiring for ranks that live in ONE process (one host thread per GPU): hand over the peers' plan
 * handles instead of IPC blobs; peer access is enabled with cudaDeviceEnablePeerAccess. */
int accfftb_plan_attach_local_peers(accfftb_plan* plan, accfftb_plan* const* plans, int nplans);

/* Batched fields.  The reference's exchange has a `howmany` path for several fields at once
 * (fast_transpose_cuda_v2_h etc., src/transpose_gpu.cpp:4454-4726, T_Plan_gpu::execute_gpu(..., howmany), :3808-4021) that its
 * plans never use; FFTW calls the same thing a "many" plan.  A plan made with howmany > 1 owns howmany sets of exchange
 * buffers per rank (memory: howmany x the single-field plan) and accepts up to howmany fields per accfftb_execute_*_many
 * call: every stage runs for all fields back to back and the fields SHARE the inter-GPU barriers -- one synchronisation
 * per exchange for the whole batch instead of one per field.  On such a plan the full-3-D gradient / divergence
 * (accfftb_grad_slow / accfftb_divergence_slow) send their three components through one set of exchanges as well.
 * Everything else about the plan is unchanged (single-field calls use field set 0). */
accfftb_plan* accfftb_plan_create_many(const int n[3], const int c_dims[2], int rank, int kind, int prec, unsigned flags, int howmany);
int accfftb_plan_howmany(const accfftb_plan* plan);

/* CUDA stream (cudaStream_t) the plan's kernels are enqueued on; default: the legacy default stream,
 * as in the reference.  blocking = 1 (default) makes every execute / operator call return only when
 * the device has finished, like the reference (cudaDeviceSynchronize, src/accfft_gpu.cpp:620);
 * blocking = 0 only enqueues. */
int accfftb_plan_set_stream(accfftb_plan* plan, void* cuda_stream);
int accfftb_plan_set_blocking(accfftb_plan* plan, int blocking);

/* Per-launch device timing: while on, every stage / barrier kernel is bracketed by a CUDA event pair on
 * the plan's stream (no host synchronisation).  _profile_read synchronises the stream, returns the summed
 * milliseconds and launch counts per slot since the last read, and resets.  Slots: 0 forward Z pass,
 * 1 forward Y, 2 forward X, 3 inverse X, 4 inverse Y, 5 inverse Z, 6 inter-GPU barrier, 7 real copy / exchange
 * stages of the single-axis paths, 8 single-axis transform / derivative stages, 9 copy-engine transfers of the slab
 * exchange (event pairs on the copy lanes; transfers on different lanes overlap, so their sum can exceed wall time).
 * (The reference only offers the coarse timer[5] of doc/timing.md.) */
#define ACCFFTB_PROFILE_SLOTS 10
int accfftb_plan_set_profiling(accfftb_plan* plan, int on);
int accfftb_plan_profile_read(accfftb_plan* plan, double ms[ACCFFTB_PROFILE_SLOTS], long long count[ACCFFTB_PROFILE_SLOTS]);

/* Which kernel serves a strided stage (slot 1 forward Y, 2 forward X, 3 inverse X, 4 inverse Y): 1 = plain kernel,
 * 2 = TMA-pipelined persistent kernel, 3 = plain kernel in its wide CTA shape (up to 1024 threads: 128-byte row segments
 * for lengths / precisions whose default shape cannot reach them; used for every stage that stores to peers, falls back to
 * 1 where the length has no wide shape), 0 = decided per launch by the size heuristic (plans made with ACCFFTB_ESTIMATE).
 * Plans made with ACCFFTB_MEASURE / ACCFFTB_PATIENT time both at creation -- on plan-owned scratch, never on user
 * data (the reference's tuner, src/transpose_gpu.cpp:632-743, runs on the caller's buffers) -- and keep the faster. */
int accfftb_plan_stage_variant(const accfftb_plan* plan, int slot);

/* The fused local pair.  On a 1 x 1 or slab (P1 == 1) grid the Z and Y passes of a transform are both local to the rank
 * (the reference's cufftExecD2Z + per-slab cufftExecZ2Z pair, src/accfft_gpu.cpp:477-516, and its mirror :560-602).  For the
 * length pairs it is compiled for (csrc/fused_combos.def) they run as ONE persistent kernel whose intermediate array is a ring
 * of a few planes that stays in L2 and never reaches HBM; the launch is profiled in the slot of its first pass (0 forward,
 * 4 inverse).  dir: 0 forward, 1 inverse.  info (may be NULL) receives what the last such launch used:
 * {planes per chunk, lag, ring slots, chunks, ring planes (0 = whole buffer), grid, threads, shared memory bytes}.
 * Returns 1 if the last transform in that direction ran fused, 0 if not.  The fused kernel is OFF by default
 * (ACCFFT_B200_FUSED=1 at plan creation switches it on): it removes the HBM traffic of the intermediate array as designed
 * (ncu: 2.2 GB instead of 4.3 GB per direction at 512^3 double) but its tiles run slower inside the persistent loop than as
 * stand-alone kernels -- 0.94 ms vs 0.74 ms per direction; DESIGN.md section 4 has the measurements. */
int accfftb_plan_fused_info(const accfftb_plan* plan, int dir, int info[8]);

long long accfftb_plan_alloc_max(const accfftb_plan* plan);
/* number of kernels the plan has launched so far (stage kernels + inter-GPU barrier kernels) */
long long accfftb_plan_launch_count(const accfftb_plan* plan);
/* number of copy-engine transfers (cudaMemcpy[2D]Async between exchange buffers) the plan has enqueued so far */
long long accfftb_plan_copy_count(const accfftb_plan* plan);

/* How the plan exchanges data between GPUs.  Returns 1 for the copy-engine exchange -- slab grids (P1 == 1): the exchanging
 * FFT stage packs into a local staging buffer, chunk by chunk, and asynchronous device-to-device copies over NVLink move
 * chunk c while chunk c+1 computes (the reference's closest relatives: fast_transpose_cuda_v1_2_h,
 * src/transpose_gpu.cpp:2643-2673, and par::Mpi_Alltoallv_dense_gpu, include/parUtils.txx:165-384, both host staged) --
 * and 0 for direct peer stores fused into the FFT kernels (pencil grids, single-rank plans, ACCFFT_B200_CE=0).
 * chunks / lanes (may be NULL) receive the chunk count and the number of concurrent copy streams. */
int accfftb_plan_exchange_mode(const accfftb_plan* plan, int* chunks, int* lanes);

/* Health of a multi-GPU plan, WITHOUT synchronising anything: 0, or ACCFFTB_ERR_TIMEOUT once an inter-GPU barrier of the
 * plan has timed out (a peer died or issued a different sequence of calls).  The condition is sticky: the stages after
 * that barrier ran on incomplete data, so every later call on the plan fails with the same code before launching
 * anything -- in blocking and non-blocking mode alike.  Callers that drive the plan with blocking = 0 should poll this
 * after their own stream synchronisation.  Time-out: 10 s of SM clock (ACCFFT_B200_BARRIER_TIMEOUT_MS overrides). */
int accfftb_plan_check(accfftb_plan* plan);

/* accfft_destroy_plan_gpu / accfft_destroy_plan (src/accfft_gpu.cpp:1150-1192); frees everything the
 * plan owns (the reference leaks the plan object and T_plan_x/xi).  NULL is accepted. */
int accfftb_plan_destroy(accfftb_plan* plan);

/* ---- execution (device pointers) ---------------------------------------------------------------
 * timer: NULL or double[5], accumulated (+=) with the reference's meaning (doc/timing.md):
 *   [0] exchange total, [1] pack, [2] communication, [3] unpack, [4] local FFT, all in seconds.
 *   Pack and unpack are fused into the FFT kernels here, so [1] = [3] = 0 and [0] = [2] = time spent in
 *   the inter-GPU barriers; [4] = time in the stage kernels (which includes their NVLink stores).
 * xyz: bit 2 = Z, bit 1 = Y, bit 0 = X like the std::bitset<3> of the reference
 *   (src/accfft_gpu.cpp:477,499,527).  An axis whose bit is clear is not transformed; its stage still moves the
 *   data into the next layout (the reference always runs its transposes), so the output layout never changes.
 *   r2c / c2r need bit 2: the real <-> complex step is the Z stage (the reference reinterprets the real rows as
 *   complex when Z is deselected, which is not a transform) -- otherwise ACCFFTB_ERR_UNSUPPORTED.
 */

/* accfft_execute_r2c_gpu / _gpuf (src/accfft_gpu.cpp:635-649): real [isize] (rows of N2 reals, or of
 * 2*(N2/2+1) when data_out == data, i.e. in place) -> complex [osize]. */
int accfftb_execute_r2c(accfftb_plan* plan, const void* data, void* data_out, double* timer, int xyz);
/* accfft_execute_c2r_gpu / _gpuf (:661-675): complex [osize] -> real [isize].  Unlike the reference
 * (:544-545) the input spectrum is NOT overwritten. */
int accfftb_execute_c2r(accfftb_plan* plan, const void* data, void* data_out, double* timer, int xyz);
/* accfft_execute_c2c_gpu / _gpuf (:943-1136): direction = ACCFFTB_FORWARD / ACCFFTB_BACKWARD. */
int accfftb_execute_c2c(accfftb_plan* plan, int direction, const void* data, void* data_out, double* timer, int xyz);

/* Batched executes on a plan made with accfftb_plan_create_many: howmany (<= the plan's field count) independent fields,
 * data[f] -> data_out[f], all axes, out of place.  Result = howmany single-field executes; the stages of all fields share
 * the inter-GPU barriers (see accfftb_plan_create_many).  Collective like the single-field calls. */
int accfftb_execute_r2c_many(accfftb_plan* plan, int howmany, const void* const* data, void* const* data_out, double* timer);
int accfftb_execute_c2r_many(accfftb_plan* plan, int howmany, const void* const* data, void* const* data_out, double* timer);
int accfftb_execute_c2c_many(accfftb_plan* plan, int direction, int howmany, const void* const* data, void* const* data_out, double* timer);

/* ---- single-axis real transforms: accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t (src/accfft_gpu.cpp:1212-1604), the
 * building blocks of the reference's per-axis gradient.  r2c plans only.  axis: 0 = x, 1 = y, 2 = z.
 * Real side: the [isize] block.  Complex side (sizes / starts: accfftb_stage_size which = 0 for z, 4 for y, 6 for x):
 *   z: [isize0][isize1][N2/2+1]                                                     (osize_0)
 *   y: [isize0][N1/2+1][n2b],  n2b = this rank's block of N2 over the row group     (osize_yi; starts: ostart_y)
 *   x: [N0/2+1][mb],  mb = this rank's block of the flattened local plane isize1*N2 over the column group (osize_xi)
 * Unnormalised, like the reference (c2r(r2c(x)) = N_axis * x).  Neither call overwrites its input (the reference's
 * c2r_x transforms in place in `data`, :1535-1538). */
int accfftb_execute_r2c_axis(accfftb_plan* plan, int axis, const void* data, void* data_out, double* timer);
int accfftb_execute_c2r_axis(accfftb_plan* plan, int axis, const void* data, void* data_out, double* timer);

/* ---- spectral operators (accfft_operators_gpu.h:44-67; src/operators_gpu.txx; kernels
 * src/operators_cuda.cu:15-215).  Real [isize] in, real [isize] out, normalised, Nyquist modes zeroed.
 * The symbol multiply and the 1/N scale are fused into the X-pass stores (no Hadamard kernels, no
 * per-call cudaMalloc).  Need an r2c plan.  timer: NULL or double[7] (the reference's operators write
 * index 6). */
int accfftb_laplace(accfftb_plan* plan, const void* A, void* LA, double* timer);
int accfftb_biharmonic(accfftb_plan* plan, const void* A, void* BA, double* timer);
/* Solver-style operators: forward transform -> symbol -> inverse transform in one call, the symbol
 * fused into the X-pass stores like laplace (zero extra passes over the spectrum).  Wave numbers follow the same
 * rule as every operator here (k = X > N/2 ? X - N : X, zero at X == N/2).
 *   accfftb_inv_laplace    : u = F^-1[ -F[A] / |k|^2 ], zero mode -> 0.  GPU twin of the reference's CPU-only
 *                            accfft_inv_laplace (include/accfft_operators.h:57-58, src/operators.txx:275-333)
 *   accfftb_inv_biharmonic : u = F^-1[  F[A] / |k|^4 ]                   (accfft_inv_biharmonic, :59-60, :335-392)
 *   accfftb_symbol_solve   : u = F^-1[ (alpha + beta |k|^2)^(+1 | -1) F[A] ]  (inverse != 0: divide; modes where the
 *                            denominator vanishes -> 0).  Covers the Poisson filter of steps/step4/kernels.cu:22-112
 *                            (alpha = 0, beta = -4 pi^2 / L^2, inverse), the explicit heat step of steps/step6
 *                            (alpha = 1, beta = -dt, direct) and implicit / Helmholtz solves (alpha = 1, beta = dt, inverse). */
int accfftb_inv_laplace(accfftb_plan* plan, const void* A, void* invLA, double* timer);
int accfftb_inv_biharmonic(accfftb_plan* plan, const void* A, void* invBA, double* timer);
int accfftb_symbol_solve(accfftb_plan* plan, const void* A, void* out, double alpha, double beta, int inverse, double* timer);
/* xyz selects components: bit 0 = x (Ax), bit 1 = y (Ay), bit 2 = z (Az), as grad_gpu's std::bitset<3>*
 * (src/operators_gpu.txx:172-278); unselected outputs may be NULL. */
int accfftb_grad(accfftb_plan* plan, const void* A, void* Ax, void* Ay, void* Az, int xyz, double* timer);
int accfftb_divergence(accfftb_plan* plan, const void* Ax, const void* Ay, const void* Az, void* div, double* timer);
/* accfftb_grad / accfftb_divergence follow the reference's default per-axis scheme (accfft_grad_gpu_t,
 * src/operators_gpu.txx:172-278: a 1-D real transform along the component's axis only), with forward transform,
 * i*k multiply and inverse transform fused into ONE kernel per component for x and y (one read + one write of the
 * field) and two passes for z.  The _slow variants are the reference's accfft_grad_gpu_slow_t /
 * accfft_divergence_gpu_slow_t (:72-170, :448-553): full 3-D transforms (1 forward + 1 inverse per component). */
int accfftb_grad_slow(accfftb_plan* plan, const void* A, void* Ax, void* Ay, void* Az, int xyz, double* timer);
int accfftb_divergence_slow(accfftb_plan* plan, const void* Ax, const void* Ay, const void* Az, void* div, double* timer);

#ifdef __cplusplus
}
#endif
#endif /* ACCFFT_B200_H */
