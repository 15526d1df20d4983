// accfft_steps.cpp -- acceptance program written against the AccFFT GPU API exactly the way the reference's step
// programs use it (steps/step1/step1_gpu.cpp: r2c/c2r round trip; steps/step2: in place; steps/step3: c2c;
// steps/step5/step5_gpu.cpp: grad / laplace / divergence / biharmonic on an analytic field), with the reference's own
// acceptance rule: relative L1 error below 1e-10 (double) -- steps/step1/step1_gpu.cpp:98-115.
// Built against include/accfft_gpu.h + libaccfft_gpu.so; run as   accfft_mpirun -n P ./accfft_steps N0 N1 N2
// Exit code 0 = every check passed on every rank.
#include <accfft_gpu.h>
#include <accfft_gpuf.h>
#include <accfft_operators_gpu.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static int g_fail = 0;
static int procid = 0;

static void verdict(const char* what, double err, double tol) {
  double worst = 0;
  MPI_Allreduce(&err, &worst, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
  const bool ok = worst <= tol && worst == worst;
  if (procid == 0) printf("  %-46s max relative error %.3e  %s\n", what, worst, ok ? "CORRECT" : "WRONG");
  if (!ok) g_fail = 1;
}

// sum |a - b| / sum |b| over all ranks
static double rel_l1(const double* a, const double* b, size_t n) {
  double loc[2] = {0, 0}, glob[2] = {0, 0};
  for (size_t i = 0; i < n; ++i) {
    loc[0] += fabs(a[i] - b[i]);
    loc[1] += fabs(b[i]);
  }
  MPI_Allreduce(loc, glob, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  return glob[1] > 0 ? glob[0] / glob[1] : glob[0];
}

// smooth periodic test field with known derivatives: f = sin(a x) cos(b y) sin(c z), x = 2 pi i / N
struct Field {
  int a, b, c;
  double f(double x, double y, double z) const { return sin(a * x) * cos(b * y) * sin(c * z); }
  double fx(double x, double y, double z) const { return a * cos(a * x) * cos(b * y) * sin(c * z); }
  double fy(double x, double y, double z) const { return -b * sin(a * x) * sin(b * y) * sin(c * z); }
  double fz(double x, double y, double z) const { return c * sin(a * x) * cos(b * y) * cos(c * z); }
  double lap(double x, double y, double z) const { return -(double)(a * a + b * b + c * c) * f(x, y, z); }
};

template <class F>
static void fill(std::vector<double>& v, const int* n, const int* isize, const int* istart, int rowlen, F fn) {
  const double pi2 = 6.283185307179586476925286766559;
  for (int i = 0; i < isize[0]; ++i)
    for (int j = 0; j < isize[1]; ++j)
      for (int k = 0; k < isize[2]; ++k)
        v[((size_t)i * isize[1] + j) * rowlen + k] =
            fn(pi2 * (i + istart[0]) / n[0], pi2 * (j + istart[1]) / n[1], pi2 * (k + istart[2]) / n[2]);
}

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int nprocs;
  MPI_Comm_rank(MPI_COMM_WORLD, &procid);
  MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
  int n[3] = {64, 64, 64};
  if (argc >= 4) {
    n[0] = atoi(argv[1]);
    n[1] = atoi(argv[2]);
    n[2] = atoi(argv[3]);
  }
  int c_dims[2] = {0, 0};   // like the steps: let accfft_create_comm pick the grid
  if (argc >= 6) {
    c_dims[0] = atoi(argv[4]);
    c_dims[1] = atoi(argv[5]);
  }
  MPI_Comm c_comm;
  accfft_create_comm(MPI_COMM_WORLD, c_dims, &c_comm);
  accfft_init();
  if (procid == 0) printf("accfft_steps: %d x %d x %d on a %d x %d process grid\n", n[0], n[1], n[2], c_dims[0], c_dims[1]);
  const double N = (double)n[0] * n[1] * n[2];
  const Field F = {2, 3, 1};

  // ---- step 1: out-of-place r2c + c2r ------------------------------------------------------------------------
  {
    int isize[3], istart[3], osize[3], ostart[3];
    const int alloc_max = accfft_local_size_dft_r2c_gpu(n, isize, istart, osize, ostart, c_comm);
    const size_t cnt = (size_t)isize[0] * isize[1] * isize[2];
    std::vector<double> host(cnt > 0 ? cnt : 1), back(cnt > 0 ? cnt : 1);
    fill(host, n, isize, istart, isize[2], [&](double x, double y, double z) { return F.f(x, y, z); });
    double *data, *data2;
    Complex* hat;
    cudaMalloc((void**)&data, (cnt + 1) * sizeof(double));
    cudaMalloc((void**)&data2, (cnt + 1) * sizeof(double));
    cudaMalloc((void**)&hat, alloc_max);
    accfft_plan_gpu* plan = accfft_plan_dft_3d_r2c_gpu(n, data, (double*)hat, c_comm, ACCFFT_MEASURE);
    if (!plan) {
      if (procid == 0) printf("plan creation failed\n");
      MPI_Finalize();
      return 3;
    }
    cudaMemcpy(data, host.data(), cnt * sizeof(double), cudaMemcpyHostToDevice);
    double timer[5] = {0};
    accfft_execute_r2c_gpu(plan, data, hat, timer);
    accfft_execute_c2r_gpu(plan, hat, data2, timer);
    cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < cnt; ++i) back[i] /= N;
    verdict("r2c + c2r round trip (step1)", rel_l1(back.data(), host.data(), cnt), 1e-10);

    // the spectrum of sin(2x) cos(3y) sin(z) has exactly 8 non-zero modes of magnitude N/8 (4 in the r2c half)
    {
      const size_t ocnt = (size_t)osize[0] * osize[1] * osize[2];
      std::vector<double> spec(2 * (ocnt > 0 ? ocnt : 1));
      cudaMemcpy(spec.data(), hat, ocnt * 2 * sizeof(double), cudaMemcpyDeviceToHost);
      double err = 0;
      for (int i = 0; i < osize[0]; ++i)
        for (int j = 0; j < osize[1]; ++j)
          for (int k = 0; k < osize[2]; ++k) {
            const int X = i + ostart[0], Y = j + ostart[1], Z = k + ostart[2];
            const bool mode = (X == F.a || X == n[0] - F.a) && (Y == F.b || Y == n[1] - F.b) && Z == F.c;
            const size_t at = ((size_t)i * osize[1] + j) * osize[2] + k;
            const double mag = hypot(spec[2 * at], spec[2 * at + 1]);
            err = fmax(err, fabs(mag - (mode ? N / 8 : 0.0)) / (N / 8));
          }
      verdict("forward spectrum of the analytic field", err, 1e-12);
    }

    // ---- step 5: spectral operators -----------------------------------------------------------------------------
    double *gx, *gy, *gz;
    cudaMalloc((void**)&gx, (cnt + 1) * sizeof(double));
    cudaMalloc((void**)&gy, (cnt + 1) * sizeof(double));
    cudaMalloc((void**)&gz, (cnt + 1) * sizeof(double));
    std::vector<double> want(cnt > 0 ? cnt : 1);
    std::bitset<3> XYZ = 7;
    accfft_grad_gpu(gx, gy, gz, data, plan, &XYZ, timer);
    double* comps[3] = {gx, gy, gz};
    const char* names[3] = {"grad x (step5)", "grad y (step5)", "grad z (step5)"};
    for (int ax = 0; ax < 3; ++ax) {
      fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) {
        return ax == 0 ? F.fx(x, y, z) : ax == 1 ? F.fy(x, y, z) : F.fz(x, y, z);
      });
      cudaMemcpy(back.data(), comps[ax], cnt * sizeof(double), cudaMemcpyDeviceToHost);
      verdict(names[ax], rel_l1(back.data(), want.data(), cnt), 1e-10);
    }
    accfft_laplace_gpu(data2, data, plan, timer);
    fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) { return F.lap(x, y, z); });
    cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    verdict("laplace (step5)", rel_l1(back.data(), want.data(), cnt), 1e-10);
    accfft_divergence_gpu(data2, gx, gy, gz, plan, timer);   // div grad f = laplace f
    cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    verdict("divergence of the gradient (step5)", rel_l1(back.data(), want.data(), cnt), 1e-10);
    accfft_biharmonic_gpu(data2, data, plan, timer);
    const double s = (double)(F.a * F.a + F.b * F.b + F.c * F.c);
    fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) { return s * s * F.f(x, y, z); });
    cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    verdict("biharmonic (step5)", rel_l1(back.data(), want.data(), cnt), 1e-10);

    // ---- the reference's other operator scheme (full 3-D transforms) and the inverse operators ----------------------
    accfft_grad_gpu_slow_t<double, accfft_plan_gpu>(gx, gy, gz, data, plan, &XYZ, timer);
    fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) { return F.fy(x, y, z); });
    cudaMemcpy(back.data(), gy, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    verdict("grad y, slow scheme (accfft_grad_gpu_slow_t)", rel_l1(back.data(), want.data(), cnt), 1e-10);
    accfft_divergence_slow_gpu(data2, gx, gy, gz, plan, timer);
    fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) { return F.lap(x, y, z); });
    cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    verdict("divergence, slow scheme", rel_l1(back.data(), want.data(), cnt), 1e-10);
    accfft_inv_laplace_gpu(data2, data, plan, timer);          // f is a single mode: inv_laplace f = -f / |k|^2
    fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) { return -F.f(x, y, z) / s; });
    cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    verdict("inverse laplace (GPU twin of accfft_inv_laplace)", rel_l1(back.data(), want.data(), cnt), 1e-10);
    accfft_symbol_solve_gpu(data2, data, 1.0, 0.25, true, plan, timer);   // implicit heat step (1 + dt |k|^2)^-1, dt = 0.25
    fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) { return F.f(x, y, z) / (1.0 + 0.25 * s); });
    cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    verdict("implicit heat step (accfft_symbol_solve_gpu)", rel_l1(back.data(), want.data(), cnt), 1e-10);

    // ---- single-axis transforms (accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t): round trip = N_axis * f ------------------
    {
      const char* an[3] = {"x", "y", "z"};
      for (int ax = 0; ax < 3; ++ax) {
        if (ax == 0) {
          accfft_execute_r2c_x_gpu_t<double, Complex>(plan, data, hat, timer);
          accfft_execute_c2r_x_gpu_t<Complex, double>(plan, hat, data2, timer);
        } else if (ax == 1) {
          accfft_execute_r2c_y_gpu_t<double, Complex>(plan, data, hat, timer);
          accfft_execute_c2r_y_gpu_t<Complex, double>(plan, hat, data2, timer);
        } else {
          accfft_execute_r2c_z_gpu_t<double, Complex>(plan, data, hat, timer);
          accfft_execute_c2r_z_gpu_t<Complex, double>(plan, hat, data2, timer);
        }
        cudaMemcpy(back.data(), data2, cnt * sizeof(double), cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < cnt; ++i) back[i] /= n[ax];
        char what[64];
        snprintf(what, sizeof(what), "single-axis r2c_%s + c2r_%s round trip", an[ax], an[ax]);
        verdict(what, rel_l1(back.data(), host.data(), cnt), 1e-10);
      }
    }
    accfft_destroy_plan_gpu(plan);

    // ---- step 2: in place (rows padded to 2 * (N2/2 + 1) reals) ----------------------------------------------------
    const int row = 2 * (n[2] / 2 + 1);
    std::vector<double> padded((size_t)isize[0] * isize[1] * row + 1, 0.0), pback(padded.size());
    fill(padded, n, isize, istart, row, [&](double x, double y, double z) { return F.f(x, y, z); });
    accfft_plan_gpu* iplan = accfft_plan_dft_3d_r2c_gpu(n, (double*)hat, (double*)hat, c_comm, ACCFFT_ESTIMATE);
    cudaMemcpy(hat, padded.data(), (size_t)isize[0] * isize[1] * row * sizeof(double), cudaMemcpyHostToDevice);
    accfft_execute_r2c_gpu(iplan, (double*)hat, hat);
    accfft_execute_c2r_gpu(iplan, hat, (double*)hat);
    cudaMemcpy(pback.data(), hat, (size_t)isize[0] * isize[1] * row * sizeof(double), cudaMemcpyDeviceToHost);
    double err = 0, ref = 0;
    for (int i = 0; i < isize[0] * isize[1]; ++i)
      for (int k = 0; k < n[2]; ++k) {
        err += fabs(pback[(size_t)i * row + k] / N - padded[(size_t)i * row + k]);
        ref += fabs(padded[(size_t)i * row + k]);
      }
    double e2[2] = {err, ref}, g2[2];
    MPI_Allreduce(e2, g2, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    verdict("in-place r2c + c2r round trip (step2)", g2[1] > 0 ? g2[0] / g2[1] : g2[0], 1e-10);
    accfft_destroy_plan(iplan);
    cudaFree(data); cudaFree(data2); cudaFree(hat); cudaFree(gx); cudaFree(gy); cudaFree(gz);
  }

  // ---- step 3: c2c, in place, forward + backward ---------------------------------------------------------------------
  {
    int isize[3], istart[3], osize[3], ostart[3];
    const int alloc_max = accfft_local_size_dft_c2c_gpu(n, isize, istart, osize, ostart, c_comm);
    const size_t cnt = (size_t)isize[0] * isize[1] * isize[2];
    std::vector<double> host(2 * cnt + 2), back(2 * cnt + 2);
    const double pi2 = 6.283185307179586476925286766559;
    for (int i = 0; i < isize[0]; ++i)
      for (int j = 0; j < isize[1]; ++j)
        for (int k = 0; k < isize[2]; ++k) {
          const double x = pi2 * (i + istart[0]) / n[0], y = pi2 * (j + istart[1]) / n[1], z = pi2 * (k + istart[2]) / n[2];
          const size_t at = ((size_t)i * isize[1] + j) * isize[2] + k;
          host[2 * at] = F.f(x, y, z);
          host[2 * at + 1] = cos(x) * sin(2 * z);
        }
    Complex* buf;
    cudaMalloc((void**)&buf, alloc_max);
    accfft_plan_gpu* plan = accfft_plan_dft_3d_c2c_gpu(n, buf, buf, c_comm);
    cudaMemcpy(buf, host.data(), cnt * 2 * sizeof(double), cudaMemcpyHostToDevice);
    accfft_execute_c2c_gpu(plan, ACCFFT_FORWARD, buf, buf);
    accfft_execute_c2c_gpu(plan, ACCFFT_BACKWARD, buf, buf);
    cudaMemcpy(back.data(), buf, cnt * 2 * sizeof(double), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < 2 * cnt; ++i) back[i] /= N;
    verdict("in-place c2c forward + backward (step3)", rel_l1(back.data(), host.data(), 2 * cnt), 1e-10);
    accfft_destroy_plan_gpu(plan);
    cudaFree(buf);
  }

  // ---- single precision API (accfft_gpuf.h) ------------------------------------------------------------------------
  {
    int isize[3], istart[3], osize[3], ostart[3];
    const int alloc_max = accfft_local_size_dft_r2c_gpuf(n, isize, istart, osize, ostart, c_comm);
    const size_t cnt = (size_t)isize[0] * isize[1] * isize[2];
    std::vector<double> hd(cnt + 1);
    fill(hd, n, isize, istart, isize[2], [&](double x, double y, double z) { return F.f(x, y, z); });
    std::vector<float> host(cnt + 1), back(cnt + 1);
    for (size_t i = 0; i < cnt; ++i) host[i] = (float)hd[i];
    float *data, *data2;
    Complexf* hat;
    cudaMalloc((void**)&data, (cnt + 1) * sizeof(float));
    cudaMalloc((void**)&data2, (cnt + 1) * sizeof(float));
    cudaMalloc((void**)&hat, alloc_max);
    accfft_plan_gpuf* plan = accfft_plan_dft_3d_r2c_gpuf(n, data, (float*)hat, c_comm);
    cudaMemcpy(data, host.data(), cnt * sizeof(float), cudaMemcpyHostToDevice);
    accfft_execute_r2c_gpuf(plan, data, hat);
    accfft_execute_c2r_gpuf(plan, hat, data2);
    cudaMemcpy(back.data(), data2, cnt * sizeof(float), cudaMemcpyDeviceToHost);
    std::vector<double> bd(cnt + 1);
    for (size_t i = 0; i < cnt; ++i) bd[i] = back[i] / N;
    verdict("single precision r2c + c2r round trip", rel_l1(bd.data(), hd.data(), cnt), 1e-5);
    accfft_laplace_gpuf(data2, data, plan);
    cudaMemcpy(back.data(), data2, cnt * sizeof(float), cudaMemcpyDeviceToHost);
    std::vector<double> want(cnt + 1);
    fill(want, n, isize, istart, isize[2], [&](double x, double y, double z) { return F.lap(x, y, z); });
    for (size_t i = 0; i < cnt; ++i) bd[i] = back[i];
    verdict("single precision laplace", rel_l1(bd.data(), want.data(), cnt), 1e-4);
    accfft_destroy_plan_gpu(plan);
    cudaFree(data); cudaFree(data2); cudaFree(hat);
  }

  accfft_cleanup_gpu();
  int all_fail = 0;
  MPI_Allreduce(&g_fail, &all_fail, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  if (procid == 0) printf(all_fail ? "accfft_steps: FAILED\n" : "accfft_steps: Results are CORRECT\n");
  MPI_Comm_free(&c_comm);
  MPI_Finalize();
  return all_fail;
}
