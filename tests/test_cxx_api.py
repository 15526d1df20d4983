"""The AccFFT-compatible C++ layer (include/accfft_gpu.h, accfft_gpuf.h, accfft_operators_gpu.h -> libaccfft_gpu.so)
and the single-node MPI shim it is built against in this MPI-less image.

CPU part: the acceptance program compiles and links against the in-tree headers / libraries exactly like a reference
step program would (`#include <accfft_gpu.h>`), libaccfft_gpu.so exports the C++ names the reference's step programs
link against, and the shim's collectives work across processes.
GPU part: the acceptance program (r2c/c2r, in place, c2c, single precision, grad / laplace / divergence / biharmonic with
the reference's own 1e-10 relative-L1 rule, steps/step1/step1_gpu.cpp:98-115) runs as 1 rank and, through
accfft_mpirun, as 2 ranks."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = os.path.join(ROOT, "tests", "cxx")
MPIRUN = os.path.join(ROOT, "accfft_b200", "bin", "accfft_mpirun")
LIB = os.path.join(ROOT, "accfft_b200", "libaccfft_gpu.so")


@pytest.fixture(scope="module")
def binaries():
    if not (os.path.exists(LIB) and os.path.exists(MPIRUN)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "accfft_b200", "csrc"), "-j", "8"])
    subprocess.check_call(["make", "-C", CXX])
    return os.path.join(CXX, "bin")


def test_acceptance_program_builds_against_the_headers(binaries):
    assert os.path.exists(os.path.join(binaries, "accfft_steps"))


def test_library_exports_the_reference_cxx_names(binaries):
    out = subprocess.check_output(["nm", "-DC", "--defined-only", LIB], text=True)
    wanted = [
        "accfft_create_comm(int, int*, int*)", "accfft_init()", "accfft_alloc(long)", "accfft_free(void*)",
        "accfft_local_size_dft_r2c_gpu(int*, int*, int*, int*, int*, int)",
        "accfft_local_size_dft_c2c_gpu(int*, int*, int*, int*, int*, int)",
        "accfft_local_size_dft_r2c_gpuf(int*, int*, int*, int*, int*, int)",
        "accfft_local_size_dft_c2c_gpuf(int*, int*, int*, int*, int*, int)",
        "accfft_plan_dft_3d_r2c_gpu(int*, double*, double*, int, unsigned int)",
        "accfft_plan_dft_3d_c2c_gpu(int*, double (*) [2], double (*) [2], int, unsigned int)",
        "accfft_plan_dft_3d_r2c_gpuf(int*, float*, float*, int, unsigned int)",
        "accfft_plan_dft_3d_c2c_gpuf(int*, float (*) [2], float (*) [2], int, unsigned int)",
        "accfft_execute_r2c_gpu(accfft_plan_gpu*, double*, double (*) [2], double*, std::bitset<3ul>)",
        "accfft_execute_c2r_gpu(accfft_plan_gpu*, double (*) [2], double*, double*, std::bitset<3ul>)",
        "accfft_execute_c2c_gpu(accfft_plan_gpu*, int, double (*) [2], double (*) [2], double*, std::bitset<3ul>)",
        "accfft_execute_gpu(accfft_plan_gpu*, int, double*, double*, double*, std::bitset<3ul>)",
        "accfft_execute_r2c_gpuf(accfft_plan_gpuf*, float*, float (*) [2], double*, std::bitset<3ul>)",
        "accfft_execute_c2r_gpuf(accfft_plan_gpuf*, float (*) [2], float*, double*, std::bitset<3ul>)",
        "accfft_execute_c2c_gpuf(accfft_plan_gpuf*, int, float (*) [2], float (*) [2], double*, std::bitset<3ul>)",
        "accfft_destroy_plan(accfft_plan_gpu*)", "accfft_destroy_plan(accfft_plan_gpuf*)",
        "accfft_destroy_plan_gpu(accfft_plan_gpu*)", "accfft_destroy_plan_gpu(accfft_plan_gpuf*)",
        "accfft_cleanup_gpu()", "accfft_cleanup_gpuf()",
        "accfft_grad_gpu(double*, double*, double*, double*, accfft_plan_gpu*, std::bitset<3ul>*, double*)",
        "accfft_laplace_gpu(double*, double*, accfft_plan_gpu*, double*)",
        "accfft_divergence_gpu(double*, double*, double*, double*, accfft_plan_gpu*, double*)",
        "accfft_biharmonic_gpu(double*, double*, accfft_plan_gpu*, double*)",
        "accfft_grad_gpuf(float*, float*, float*, float*, accfft_plan_gpuf*, std::bitset<3ul>*, double*)",
        "accfft_laplace_gpuf(float*, float*, accfft_plan_gpuf*, double*)",
        "accfft_divergence_gpuf(float*, float*, float*, float*, accfft_plan_gpuf*, double*)",
        "accfft_biharmonic_gpuf(float*, float*, accfft_plan_gpuf*, double*)",
        # single-axis executes and operator templates the reference's operators TU links against
        "void accfft_execute_r2c_x_gpu_t<double, double [2]>(accfft_plan_gpu*, double*, double (*) [2], double*)",
        "void accfft_execute_c2r_x_gpu_t<double [2], double>(accfft_plan_gpu*, double (*) [2], double*, double*)",
        "void accfft_execute_r2c_y_gpu_t<double, double [2]>(accfft_plan_gpu*, double*, double (*) [2], double*)",
        "void accfft_execute_c2r_y_gpu_t<double [2], double>(accfft_plan_gpu*, double (*) [2], double*, double*)",
        "void accfft_execute_r2c_z_gpu_t<double, double [2]>(accfft_plan_gpu*, double*, double (*) [2], double*)",
        "void accfft_execute_c2r_z_gpu_t<double [2], double>(accfft_plan_gpu*, double (*) [2], double*, double*)",
        "void accfft_execute_r2c_y_gpu_t<float, float [2]>(accfft_plan_gpuf*, float*, float (*) [2], double*)",
        "void accfft_grad_gpu_slow_t<double, accfft_plan_gpu>(double*, double*, double*, double*, accfft_plan_gpu*, std::bitset<3ul>*, double*)",
        "accfft_divergence_slow_gpu(double*, double*, double*, double*, accfft_plan_gpu*, double*)",
        # GPU twins of the reference's CPU-only inverse operators (extensions)
        "accfft_inv_laplace_gpu(double*, double*, accfft_plan_gpu*, double*)",
        "accfft_inv_biharmonic_gpu(double*, double*, accfft_plan_gpu*, double*)",
    ]
    missing = [w for w in wanted if w not in out]
    assert not missing, missing


@pytest.mark.parametrize("ranks", [1, 2, 4, 8])
def test_mpi_shim_collectives(binaries, ranks):
    exe = os.path.join(binaries, "mpi_shim_check")
    cmd = [exe] if ranks == 1 else [MPIRUN, "-n", str(ranks), "--gpus-per-node", "0", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("n,ranks", [((12, 10, 14), 1), ((12, 10, 14), 4), ((5, 7, 6), 8), ((16, 3, 9), 2)])
def test_npy_io_neighbour(binaries, tmp_path, n, ranks):
    """write_npy / read_npy (include/accfft_utils.h): the .npy stand-in for the reference's read_pnetcdf / write_pnetcdf
    (include/accfft_utils.h:24-38).  Ranks write their input pencils into ONE global file, read other pencils back, and numpy
    sees the global C-order array.  CPU only (the shim runs without GPUs)."""
    import numpy as np
    exe = os.path.join(binaries, "npy_io_check")
    path = str(tmp_path / "field.npy")
    args = [path] + [str(v) for v in n]
    cmd = [exe] + args if ranks == 1 else [MPIRUN, "-n", str(ranks), "--gpus-per-node", "0", exe] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    a = np.load(path)
    assert a.shape == tuple(n) and a.dtype == np.float64
    assert np.array_equal(a.ravel(), np.arange(n[0] * n[1] * n[2], dtype=np.float64))


@pytest.mark.gpu
@pytest.mark.parametrize("n,ranks", [((64, 64, 64), 1), ((32, 48, 40), 1), ((64, 64, 64), 2), ((24, 20, 36), 2)])
def test_acceptance_program_on_gpu(binaries, n, ranks):
    """ranks = 2: two PROCESSES share cuda:0 (the exchange then runs through CUDA IPC mappings of the peers' slabs,
    the same code path as one process per GPU)."""
    exe = os.path.join(binaries, "accfft_steps")
    args = [str(v) for v in n]
    cmd = [exe] + args if ranks == 1 else [MPIRUN, "-n", str(ranks), "--gpus-per-node", "1", exe] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Results are CORRECT" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
