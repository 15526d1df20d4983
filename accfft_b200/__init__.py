"""accfft_b200 -- B200-native distributed 3D FFT engine behind AccFFT's GPU API.

The compute path is libaccfft_b200.so (hand-written sm_100a kernels, C ABI in include/accfft_b200.h).
This package is the thin ctypes host layer that mirrors the reference's operator/plan interface
(include/accfft_gpu.h, accfft_gpuf.h, accfft_operators_gpu.h of AccFFT) for Python callers, the tests
and the benchmark.  There is NO CPU fallback: importing works without a GPU (size queries are pure
integer math), every compute call requires the CUDA library and a device.
"""
from .api import (  # noqa: F401
    ACCFFT_BACKWARD,
    ACCFFT_ESTIMATE,
    ACCFFT_FORWARD,
    ACCFFT_MEASURE,
    AccfftError,
    Comm,
    Plan,
    accfft_create_comm,
    attach_local_peers,
    accfft_local_size_dft_c2c_gpu,
    accfft_local_size_dft_c2c_gpuf,
    accfft_local_size_dft_r2c_gpu,
    accfft_local_size_dft_r2c_gpuf,
    accfft_plan_dft_3d_c2c_gpu,
    accfft_plan_dft_3d_c2c_gpuf,
    accfft_plan_dft_3d_r2c_gpu,
    accfft_plan_dft_3d_r2c_gpuf,
    length_fast,
    length_supported,
    lib,
    lib_path,
)
