// aux.cu -- small N-sized passes around the hot path (HBM-bound, one launch each):
//   rlvi_sigmoid_f64       standard-learning/utils.py:7-16   overflow-free logistic function of an N-vector
//   rlvi_online_ce_f64     online-learning/main.py:84-85     -t l - (1 - t) l, the label-independent "cross-entropy"
//   rlvi_irls_weights_f64  the Newton / IRLS curvature weights h_i = pi_i s_i (1 - s_i) of the L2-regularised logistic
//                          objective that utils.py:61-73 hands to liblinear (s_i or 1 - s_i = exp(-cross-entropy_i))
//   rlvi_rrm_sum_f64       standard-learning/rrm.py:12-33 and online-learning/main.py:61-81 (update_weights_rrm): the
//                          sum over max(exp(-l_i / alpha), cutoff) that SciPy's Brent evaluates, and the final weights
#include <math.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sigmoid_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double v = x[i];
    const double z = exp(-fabs(v));                       // utils.py:12-15
    out[i] = (v >= 0.0 ? 1.0 : z) / (1.0 + z);
  }
}

__global__ void __launch_bounds__(256) online_ce_kernel(const double* __restrict__ lp, const double* __restrict__ t, int64_t n,
                                                        double* __restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double ti = t[i], li = lp[i];
    out[i] = __dsub_rn(__dmul_rn(-ti, li), __dmul_rn(1.0 - ti, li));   // main.py:85, term by term (no FMA contraction)
  }
}

__global__ void __launch_bounds__(256) irls_weights_kernel(const double* __restrict__ e, const double* __restrict__ pi,
                                                           int64_t n, double* __restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double s = e[i];
    out[i] = pi[i] * s * (1.0 - s);
  }
}

__global__ void __launch_bounds__(256) rrm_sum_kernel(const double* __restrict__ losses, int64_t n, double inv_alpha,
                                                      double cutoff, double norm, double* w_out, double* partials,
                                                      unsigned int* ticket, double* out_sum) {
  __shared__ double s_red[8];
  double acc = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double phi = exp(-losses[i] * inv_alpha);       // rrm.py:18,27
    acc += (phi < cutoff) ? cutoff : phi;                 // rrm.py:19,28 (the clipping only enters the sum)
    if (w_out) w_out[i] = phi * norm;                     // rrm.py:32
  }
  double v[1] = {acc};
  block_sum<1>(v, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
  if (last_block_ticket(ticket, gridDim.x)) {
    if (threadIdx.x < 32) {
      double a = 0.0;
      for (unsigned int j = threadIdx.x; j < gridDim.x; j += 32) a += partials[j];
      a = warp_sum(a);
      if (threadIdx.x == 0) out_sum[0] = a;
    }
  }
}

int grid_for(const rlvi_ctx* ctx, int64_t n, int per_thread) {
  int64_t want = (n + 256 * per_thread - 1) / (256 * per_thread);
  const int64_t cap = int64_t(ctx->sm_count) * 8;
  return int(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" int rlvi_sigmoid_f64(rlvi_ctx* ctx, const double* x, int64_t n, double* out, void* stream) {
  RLVI_REQUIRE(ctx && x && out, "null pointer");
  RLVI_REQUIRE(n > 0, "n must be positive");
  RlviDeviceGuard guard(ctx->device);
  sigmoid_kernel<<<grid_for(ctx, n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}

extern "C" int rlvi_online_ce_f64(rlvi_ctx* ctx, const double* log_proba, const double* targets, int64_t n, double* out,
                                  void* stream) {
  RLVI_REQUIRE(ctx && log_proba && targets && out, "null pointer");
  RLVI_REQUIRE(n > 0, "n must be positive");
  RlviDeviceGuard guard(ctx->device);
  online_ce_kernel<<<grid_for(ctx, n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(log_proba, targets, n, out);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}

extern "C" int rlvi_irls_weights_f64(rlvi_ctx* ctx, const double* e, const double* weights, int64_t n, double* out,
                                     void* stream) {
  RLVI_REQUIRE(ctx && e && weights && out, "null pointer");
  RLVI_REQUIRE(n > 0, "n must be positive");
  RlviDeviceGuard guard(ctx->device);
  irls_weights_kernel<<<grid_for(ctx, n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(e, weights, n, out);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}

extern "C" int rlvi_rrm_sum_f64(rlvi_ctx* ctx, const double* losses, int64_t n, double inv_alpha, double cutoff,
                                double norm, double* w_out, double* out_sum, void* stream) {
  RLVI_REQUIRE(ctx && losses && out_sum, "null pointer");
  RLVI_REQUIRE(n > 0, "n must be positive");
  RlviDeviceGuard guard(ctx->device);
  const int grid = grid_for(ctx, n, 4);
  void* scratch = nullptr;
  int rc = rlvi_scratch(ctx, 4096 + size_t(grid) * sizeof(double), &scratch);
  if (rc != RLVI_OK) return rc;
  rrm_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      losses, n, inv_alpha, cutoff, norm, w_out, reinterpret_cast<double*>(static_cast<char*>(scratch) + 4096),
      reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + 128), out_sum);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}
