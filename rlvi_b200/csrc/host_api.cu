// host_api.cu -- host-buffer entry point: one E+M step of the logistic model with the H2D/D2H
// traffic inside the call (what the NumPy-facing drop-in does; bench.py's `e2e`).
//
// X arrives from the host in row chunks on a copy stream; the loss kernel for chunk k runs on the
// compute stream while chunk k+1 is still in flight (PCIe is the bound, the kernel hides under it).
// X stays resident, so the statistics pass does not cross PCIe again.
#include "common.cuh"

namespace {
constexpr int64_t kChunkRows = int64_t(1) << 19;   // 256 MiB of X per chunk at d = 64
}

static int em_step_logistic_host_impl(rlvi_ctx* ctx, const double* X_host, const double* y_host, int64_t n, int d,
                                      const double* params_host, double tol, int maxiter, double* pi_host,
                                      double* moments_host, rlvi_fp_result* result_host, const rlvi_fp_dist* fp_dist,
                                      const rlvi_fp_dist* stats_dist);

extern "C" int rlvi_em_step_logistic_host(rlvi_ctx* ctx, const double* X_host, const double* y_host, int64_t n,
                                          int d, const double* params_host, double tol, int maxiter,
                                          double* pi_host, double* moments_host, rlvi_fp_result* result_host) {
  return em_step_logistic_host_impl(ctx, X_host, y_host, n, d, params_host, tol, maxiter, pi_host, moments_host,
                                    result_host, nullptr, nullptr);
}

extern "C" int rlvi_em_step_logistic_host_sharded(rlvi_ctx* ctx, const double* X_host, const double* y_host,
                                                  int64_t n, int d, const double* params_host, double tol,
                                                  int maxiter, double* pi_host, double* moments_host,
                                                  rlvi_fp_result* result_host, const rlvi_fp_dist* fp_dist,
                                                  const rlvi_fp_dist* stats_dist) {
  RLVI_REQUIRE(fp_dist && stats_dist, "null rlvi_fp_dist");
  RLVI_REQUIRE(rlvi_moments_out_doubles(d) <= RLVI_DIST_STATS_CAPACITY, "statistics do not fit the peer window (d <= 88)");
  return em_step_logistic_host_impl(ctx, X_host, y_host, n, d, params_host, tol, maxiter, pi_host, moments_host,
                                    result_host, fp_dist, stats_dist);
}

static int em_step_logistic_host_impl(rlvi_ctx* ctx, const double* X_host, const double* y_host, int64_t n, int d,
                                      const double* params_host, double tol, int maxiter, double* pi_host,
                                      double* moments_host, rlvi_fp_result* result_host, const rlvi_fp_dist* fp_dist,
                                      const rlvi_fp_dist* stats_dist) {
  RLVI_REQUIRE(ctx && X_host && y_host && params_host && moments_host && result_host, "null pointer");
  RLVI_REQUIRE(n > 0 && d > 0, "n and d must be positive");
  RlviDeviceGuard guard(ctx->device);
  const int nm = rlvi_moments_out_doubles(d);
  // device layout: X | y | e | pi | params | moments | result   (each 256-byte aligned)
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t bX = up(size_t(n) * d * 8), bV = up(size_t(n) * 8), bP = up(size_t(d + 1) * 8), bM = up(size_t(nm) * 8);
  const size_t total = bX + 3 * bV + bP + bM + 256;
  if (total > ctx->big_bytes) {
    RLVI_CUDA(cudaDeviceSynchronize());
    if (ctx->big) cudaFree(ctx->big);
    ctx->big = nullptr;
    ctx->big_bytes = 0;
    if (cudaMalloc(&ctx->big, total) != cudaSuccess) {
      rlvi_set_error("cudaMalloc of %zu bytes for the device-resident copy failed", total);
      return RLVI_ERR_NOMEM;
    }
    ctx->big_bytes = total;
  }
  char* base = static_cast<char*>(ctx->big);
  double* dX = reinterpret_cast<double*>(base);
  double* dY = reinterpret_cast<double*>(base + bX);
  double* dE = reinterpret_cast<double*>(base + bX + bV);
  double* dPi = reinterpret_cast<double*>(base + bX + 2 * bV);
  double* dParams = reinterpret_cast<double*>(base + bX + 3 * bV);
  double* dMom = reinterpret_cast<double*>(base + bX + 3 * bV + bP);
  rlvi_fp_result* dRes = reinterpret_cast<rlvi_fp_result*>(base + bX + 3 * bV + bP + bM);

  cudaStream_t cs = ctx->copy_stream;   // copies
  cudaStream_t ks = nullptr;            // kernels: legacy default stream (copy stream is non-blocking)
  RLVI_CUDA(cudaMemcpyAsync(dParams, params_host, size_t(d + 1) * 8, cudaMemcpyHostToDevice, cs));
  int evi = 0;
  for (int64_t r0 = 0; r0 < n; r0 += kChunkRows) {
    const int64_t rows = (n - r0 < kChunkRows) ? (n - r0) : kChunkRows;
    RLVI_CUDA(cudaMemcpyAsync(dX + r0 * d, X_host + r0 * d, size_t(rows) * d * 8, cudaMemcpyHostToDevice, cs));
    RLVI_CUDA(cudaMemcpyAsync(dY + r0, y_host + r0, size_t(rows) * 8, cudaMemcpyHostToDevice, cs));
    cudaEvent_t ev = ctx->ev[evi];
    evi = (evi + 1) & 3;
    RLVI_CUDA(cudaEventRecord(ev, cs));
    RLVI_CUDA(cudaStreamWaitEvent(ks, ev, 0));
    int rc = rlvi_loss_f64(ctx, RLVI_LOSS_LOGISTIC_CE, 1, dX + r0 * d, dY + r0, rows, d, dParams, nullptr, nullptr,
                           dE + r0, nullptr, ks);
    if (rc != RLVI_OK) return rc;
  }
  int rc = rlvi_fixed_point_f64(ctx, RLVI_FP_STANDARD, nullptr, nullptr, dE, n, tol, maxiter, dPi, dRes, fp_dist, ks);
  if (rc != RLVI_OK) return rc;
  // pi goes back over PCIe on the copy stream while the statistics kernel runs on the compute stream
  if (pi_host) {
    RLVI_CUDA(cudaEventRecord(ctx->ev[0], ks));
    RLVI_CUDA(cudaStreamWaitEvent(cs, ctx->ev[0], 0));
    RLVI_CUDA(cudaMemcpyAsync(pi_host, dPi, size_t(n) * 8, cudaMemcpyDeviceToHost, cs));
  }
  rc = rlvi_weighted_moments_f64(ctx, dX, nullptr, dPi, n, d, 1, 1, dMom, ks);
  if (rc != RLVI_OK) return rc;
  if (stats_dist) {   // sharded: the statistics of all ranks, same bits everywhere
    rc = rlvi_stats_allreduce_f64(ctx, dMom, nm, stats_dist, ks);
    if (rc != RLVI_OK) return rc;
  }
  RLVI_CUDA(cudaMemcpyAsync(moments_host, dMom, size_t(nm) * 8, cudaMemcpyDeviceToHost, ks));
  RLVI_CUDA(cudaMemcpyAsync(result_host, dRes, sizeof(rlvi_fp_result), cudaMemcpyDeviceToHost, ks));
  RLVI_CUDA(cudaStreamSynchronize(ks));
  RLVI_CUDA(cudaStreamSynchronize(cs));
  return RLVI_OK;
}
