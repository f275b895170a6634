// abi_smoke.cu -- the C ABI used from plain C/C++ with nothing but the CUDA runtime (no Python, no torch):
// one E+M step of the logistic model through rlvi_loss_f64 -> rlvi_fixed_point_f64 -> rlvi_weighted_moments_f64,
// checked against a scalar restatement of the reference arithmetic (standard-learning/utils.py:19-21,
// rlvi.py:8-20, utils.py:36-38) written in this file.  Built by __graft_entry__.build(); run by
// tests/test_gpu_parity.py::test_c_abi_from_c.  Exit code 0 = parity within 1e-9.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "rlvi_b200.h"

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e_));              \
      return 2;                                                               \
    }                                                                         \
  } while (0)
#define RK(x)                                                                 \
  do {                                                                        \
    int r_ = (x);                                                             \
    if (r_ != RLVI_OK) {                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #x, r_, rlvi_last_error());           \
      return 3;                                                               \
    }                                                                         \
  } while (0)

static double lcg(unsigned long long* s) {   // uniform (0,1)
  *s = *s * 6364136223846793005ull + 1442695040888963407ull;
  return ((*s >> 11) + 0.5) / 9007199254740992.0;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 50000, d = 64;
  unsigned long long seed = 42;
  std::vector<double> X(size_t(n) * d), y(n), params(d + 1);
  for (auto& v : X) v = 2.0 * lcg(&seed) - 1.0;
  for (int j = 0; j <= d; ++j) params[j] = (lcg(&seed) - 0.5) * 0.5;
  for (int i = 0; i < n; ++i) y[i] = lcg(&seed) < 0.5 ? 1.0 : 0.0;

  // ---- scalar restatement ---------------------------------------------------------------------------
  std::vector<double> e(n), pi(n, 0.95), pn(n);
  for (int i = 0; i < n; ++i) {
    double phi = params[0];
    for (int j = 0; j < d; ++j) phi += X[size_t(i) * d + j] * params[1 + j];
    const double loss = -y[i] * phi + phi + log1p(exp(-phi));      // utils.py:21
    e[i] = exp(-loss);
  }
  int iters = 0;
  double eps = 0.0;
  for (int k = 1; k <= 100; ++k) {                                  // rlvi.py:8-20
    double mean = 0.0;
    for (int i = 0; i < n; ++i) mean += pi[i];
    mean /= n;
    eps = 1.0 - mean;
    const double rho = eps / (1.0 - eps);
    double err = 0.0;
    for (int i = 0; i < n; ++i) {
      pn[i] = e[i] / (rho + e[i]);
      err += (pn[i] - pi[i]) * (pn[i] - pi[i]);
    }
    pi = pn;
    iters = k;
    if (sqrt(err) < 1e-3) break;
  }
  std::vector<double> G(size_t(d) * d, 0.0);
  double S0 = 0.0;
  for (int i = 0; i < n; ++i) {
    S0 += pi[i];
    for (int a = 0; a < d; ++a)
      for (int b = a; b < d; ++b) G[size_t(a) * d + b] += pi[i] * X[size_t(i) * d + a] * X[size_t(i) * d + b];
  }

  // ---- the library ------------------------------------------------------------------------------------
  rlvi_ctx* ctx = nullptr;
  RK(rlvi_ctx_create(0, &ctx));
  double *dX, *dy, *dp, *de, *dpi, *dmom;
  rlvi_fp_result* dres;
  const int nm = rlvi_moments_out_doubles(d);
  CK(cudaMalloc(&dX, X.size() * 8));
  CK(cudaMalloc(&dy, size_t(n) * 8));
  CK(cudaMalloc(&dp, (d + 1) * 8));
  CK(cudaMalloc(&de, size_t(n) * 8));
  CK(cudaMalloc(&dpi, size_t(n) * 8));
  CK(cudaMalloc(&dmom, size_t(nm) * 8));
  CK(cudaMalloc(&dres, sizeof(rlvi_fp_result)));
  CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, y.data(), size_t(n) * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp, params.data(), (d + 1) * 8, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  RK(rlvi_loss_f64(ctx, RLVI_LOSS_LOGISTIC_CE, 1, dX, dy, n, d, dp, nullptr, nullptr, de, nullptr, st));
  RK(rlvi_fixed_point_f64(ctx, RLVI_FP_STANDARD, nullptr, nullptr, de, n, 1e-3, 100, dpi, dres, nullptr, st));
  RK(rlvi_weighted_moments_f64(ctx, dX, nullptr, dpi, n, d, 1, 1, dmom, st));
  CK(cudaStreamSynchronize(st));
  rlvi_fp_result res;
  std::vector<double> mom(nm), hpi(n);
  CK(cudaMemcpy(&res, dres, sizeof(res), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(mom.data(), dmom, size_t(nm) * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hpi.data(), dpi, size_t(n) * 8, cudaMemcpyDeviceToHost));

  double max_pi = 0.0, err_pi = 0.0, max_g = 0.0, err_g = 0.0;
  for (int i = 0; i < n; ++i) {
    max_pi = fmax(max_pi, fabs(pi[i]));
    err_pi = fmax(err_pi, fabs(pi[i] - hpi[i]));
  }
  for (int a = 0; a < d; ++a)
    for (int b = a; b < d; ++b) {
      const double ref = G[size_t(a) * d + b] / S0, got = mom[2 + 2 * d + size_t(a) * d + b] / mom[0];
      max_g = fmax(max_g, fabs(ref));
      err_g = fmax(err_g, fabs(ref - got));
    }
  printf("n=%d iters=%d/%d eps=%.15g/%.15g rel_err_pi=%.3e rel_err_G=%.3e launches=%lld\n", n, res.iters, iters, res.eps,
         eps, err_pi / max_pi, err_g / max_g, (long long)rlvi_ctx_launch_count(ctx));
  const bool ok = res.iters == iters && fabs(res.eps - eps) < 1e-9 && err_pi / max_pi < 1e-9 && err_g / max_g < 1e-9;
  // error path: a bad argument comes back as a code + message
  const int rc = rlvi_fixed_point_f64(ctx, RLVI_FP_STANDARD, nullptr, nullptr, de, n, 1e-3, 0, dpi, dres, nullptr, st);
  const bool ok_err = (rc == RLVI_ERR_INVALID) && rlvi_last_error()[0] != 0;
  rlvi_ctx_destroy(ctx);
  printf("%s\n", ok && ok_err ? "C-ABI PARITY OK" : "C-ABI PARITY FAILED");
  return ok && ok_err ? 0 : 1;
}
