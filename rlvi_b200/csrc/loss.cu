// loss.cu -- per-sample negative log-likelihoods: ONE streaming pass over X (HBM-bound).
//
// Replaces the NumPy expressions of
//   standard-learning/utils.py:19-21   cross_entropy            (RLVI_LOSS_LOGISTIC_CE)
//   standard-learning/utils.py:62-64   sklearn_log_reg's loss   (RLVI_LOSS_SOFTPLUS)
//   standard-learning/rlvi.py:72,81    (y - X theta)^2          (RLVI_LOSS_SQRES)
//   standard-learning/rlvi.py:49,57    ||theta - x||^2          (RLVI_LOSS_SQDIST)
//   standard-learning/utils.py:77-79   PCA reconstruction       (RLVI_LOSS_PCA)
//   standard-learning/utils.py:93-101  Gaussian NLL             (RLVI_LOSS_GAUSSIAN, DMMA kernel below)
// and fuses e_i = exp(-l_i) (the fixed point's input) and the pi-weighted loss sum
// (sigma2 = pi.r2 / sum pi, rlvi.py:50,58,73,82) into the same pass.
//
// Algorithmic bytes per sample: d*8 (X) + 8 (y) + 8 (pi, optional) + 8 (l) + 8 (e).
#include <math.h>

#include "rowmap.cuh"

namespace {

constexpr int kLossThreads = 256;

struct LossParams {
  const double* X;
  const double* y;
  const double* params;
  const double* w;
  double* losses;
  double* e_out;
  double* wsum_out;
  int64_t n;
  int d;
  int kind;
  int intercept;
  int L;
  double* partials;       // [grid][2]
  unsigned int* ticket;
};

// DOTK 0: a = x.theta          (LOGISTIC_CE, SOFTPLUS, SQRES)
// DOTK 1: a = x.theta, b = x.x (PCA)
// DOTK 2: a = ||theta - x||^2  (SQDIST)
template <int DOTK, int FPL, bool VEC>
__global__ void __launch_bounds__(kLossThreads) loss_kernel(const LossParams p) {
  extern __shared__ double sm[];
  double* sm_params = sm;                    // d + 1
  double* sm_red = sm + (p.d + 2);           // 2 * nwarps
  const int np = p.d + (p.intercept ? 1 : 0);
  for (int i = threadIdx.x; i < np; i += blockDim.x) sm_params[i] = p.params[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int L = p.L;
  const int q = lane & (L - 1);
  const int sub = lane / L;
  const int R = 32 / L;
  const double b0 = p.intercept ? sm_params[0] : 0.0;
  double th[FPL];
  RowMap<FPL, VEC>::load_vec(sm_params + (p.intercept ? 1 : 0), p.d, q, L, th);

  const int64_t warps_total = int64_t(gridDim.x) * (blockDim.x >> 5);
  const int64_t warp_id = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double s_wl = 0.0, s_w = 0.0;

  for (int64_t row0 = warp_id * R; row0 < p.n; row0 += warps_total * R) {
    const int64_t row = row0 + sub;
    const bool valid = row < p.n;
    double x[FPL];
    RowMap<FPL, VEC>::load_row(p.X, row, p.d, q, L, valid, x);
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < FPL; ++k) {
      if (DOTK == 2) {
        const double t = th[k] - x[k];
        a = fma(t, t, a);
      } else {
        a = fma(x[k], th[k], a);
        if (DOTK == 1) b = fma(x[k], x[k], b);
      }
    }
    a = group_sum(a, L);
    if (DOTK == 1) b = group_sum(b, L);
    if (q == 0 && valid) {
      double loss;
      switch (p.kind) {
        case RLVI_LOSS_LOGISTIC_CE: {
          const double phi = b0 + a;
          const double yi = p.y[row];
          loss = (-yi * phi + phi) + log1p(exp(-phi));          // utils.py:21, same operation order
          break;
        }
        case RLVI_LOSS_SOFTPLUS: {
          const double phi = b0 + a;
          loss = fmax(phi, 0.0) + log1p(exp(-fabs(phi)));        // logaddexp(0, phi)
          break;
        }
        case RLVI_LOSS_SQRES: {
          const double r = p.y[row] - (b0 + a);
          loss = r * r;
          break;
        }
        case RLVI_LOSS_SQDIST: {
          const double r = sqrt(a);                              // np.linalg.norm(...)**2
          loss = r * r;
          break;
        }
        default: {                                               // RLVI_LOSS_PCA
          loss = b - a * a;
          break;
        }
      }
      if (p.losses) p.losses[row] = loss;
      if (p.e_out) p.e_out[row] = exp(-loss);
      if (p.w) {
        const double wi = p.w[row];
        s_wl = fma(wi, loss, s_wl);
        s_w += wi;
      }
    }
  }

  if (p.w) {   // uniform across the grid
    double v[2] = {s_wl, s_w};
    block_sum<2>(v, sm_red);
    if (threadIdx.x == 0) {
      p.partials[2 * blockIdx.x] = v[0];
      p.partials[2 * blockIdx.x + 1] = v[1];
    }
    if (last_block_ticket(p.ticket, gridDim.x)) {
      // fixed-order final sum by warp 0 of the last block
      if (threadIdx.x < 32) {
        double a0 = 0.0, a1 = 0.0;
        for (unsigned int j = lane; j < gridDim.x; j += 32) {
          a0 += p.partials[2 * j];
          a1 += p.partials[2 * j + 1];
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) {
          p.wsum_out[0] = a0;
          p.wsum_out[1] = a1;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gaussian NLL through the FP64 tensor pipe.
//   l_i = 0.5 (||U (x_i - mu)||^2 + c),  U upper triangular with U^T U = cov^-1.
// A warp takes 8 rows per step.  Z = C U^T is accumulated with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4):
// A fragment = centred samples (row g, feature pair of lane t), B fragment = U rows from shared memory
// (pitch d+8 doubles: conflict-free 128-bit reads), only the block-upper-triangular part is visited:
// nb(nb+1) DMMAs per 8 rows, nb = ceil(d/8).  The k-slots of one DMMA pair are the two halves of one
// 128-bit load, i.e. features (8kk+2t, 8kk+2t+1); the same permutation is applied to U's columns, so
// the contraction is unchanged.
// ---------------------------------------------------------------------------------------------
struct GaussParams {
  const double* X;
  const double* params;   // [c, mu(d), U(d*d) row-major upper triangular]
  const double* w;
  double* losses;
  double* e_out;
  double* wsum_out;
  int64_t n;
  int d;
  double* partials;
  unsigned int* ticket;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NB_MAX, bool VEC>
__global__ void __launch_bounds__(kLossThreads) gaussian_loss_kernel(const GaussParams p) {
  extern __shared__ __align__(16) double smg[];
  const int d = p.d;
  const int nb = (d + 7) >> 3;
  const int dp = nb * 8;
  const int pitch = dp + 8;
  double* sU = smg;                       // dp * pitch
  double* sMu = sU + size_t(dp) * pitch;  // dp
  double* sRed = sMu + dp;                // 2 * nwarps
  for (int i = threadIdx.x; i < dp * pitch; i += blockDim.x) {
    const int r = i / pitch, c = i - r * pitch;
    sU[i] = (r < d && c < d && c >= r) ? p.params[1 + d + size_t(r) * d + c] : 0.0;
  }
  for (int i = threadIdx.x; i < dp; i += blockDim.x) sMu[i] = (i < d) ? p.params[1 + i] : 0.0;
  __syncthreads();
  const double cst = p.params[0];

  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t warps_total = int64_t(gridDim.x) * (blockDim.x >> 5);
  const int64_t warp_id = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double s_wl = 0.0, s_w = 0.0;

  for (int64_t row0 = warp_id * 8; row0 < p.n; row0 += warps_total * 8) {
    const int64_t row = row0 + g;
    const bool valid = row < p.n;
    double2 c[NB_MAX];
#pragma unroll
    for (int kk = 0; kk < NB_MAX; ++kk) {
      const int f = 8 * kk + 2 * t;
      double2 v = make_double2(0.0, 0.0);
      if (kk < nb && valid) {
        const double* src = p.X + row * int64_t(d) + f;
        if (VEC) {
          if (f < d) v = ld_stream_d2(src);
        } else {
          if (f < d) v.x = ld_stream_d1(src);
          if (f + 1 < d) v.y = ld_stream_d1(src + 1);
        }
        const double2 m = *reinterpret_cast<const double2*>(sMu + f);
        v.x -= m.x;
        v.y -= m.y;
        if (f >= d) v.x = 0.0;
        if (f + 1 >= d) v.y = 0.0;
      }
      c[kk] = v;
    }
    double quad = 0.0;
#pragma unroll
    for (int nbk = 0; nbk < NB_MAX; ++nbk) {
      if (nbk < nb) {
        double z0 = 0.0, z1 = 0.0;
        const double* urow = sU + size_t(8 * nbk + g) * pitch + 2 * t;
#pragma unroll
        for (int kk = nbk; kk < NB_MAX; ++kk) {
          if (kk < nb) {
            const double2 u = *reinterpret_cast<const double2*>(urow + 8 * kk);
            dmma884(z0, z1, c[kk].x, u.x);
            dmma884(z0, z1, c[kk].y, u.y);
          }
        }
        quad = fma(z0, z0, quad);
        quad = fma(z1, z1, quad);
      }
    }
    quad += __shfl_xor_sync(0xffffffffu, quad, 1);
    quad += __shfl_xor_sync(0xffffffffu, quad, 2);
    if (t == 0 && valid) {
      const double loss = 0.5 * (quad + cst);                    // utils.py:101
      if (p.losses) p.losses[row] = loss;
      if (p.e_out) p.e_out[row] = exp(-loss);
      if (p.w) {
        const double wi = p.w[row];
        s_wl = fma(wi, loss, s_wl);
        s_w += wi;
      }
    }
  }
  if (p.w) {
    double v[2] = {s_wl, s_w};
    block_sum<2>(v, sRed);
    if (threadIdx.x == 0) {
      p.partials[2 * blockIdx.x] = v[0];
      p.partials[2 * blockIdx.x + 1] = v[1];
    }
    if (last_block_ticket(p.ticket, gridDim.x)) {
      if (threadIdx.x < 32) {
        double a0 = 0.0, a1 = 0.0;
        for (unsigned int j = lane; j < gridDim.x; j += 32) {
          a0 += p.partials[2 * j];
          a1 += p.partials[2 * j + 1];
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) {
          p.wsum_out[0] = a0;
          p.wsum_out[1] = a1;
        }
      }
    }
  }
}

template <int DOTK>
int launch_loss(rlvi_ctx* ctx, const RowMapCfg& cfg, LossParams& p, int grid, size_t smem, cudaStream_t st) {
#define RLVI_LOSS_CASE(F, V)                                                            \
  if (cfg.fpl == F && cfg.vec == V) {                                                   \
    loss_kernel<DOTK, F, V><<<grid, kLossThreads, smem, st>>>(p);                       \
    RLVI_LAUNCH_CHECK(ctx);                                                             \
    return RLVI_OK;                                                                     \
  }
  RLVI_LOSS_CASE(4, true)
  RLVI_LOSS_CASE(4, false)
  RLVI_LOSS_CASE(16, true)
  RLVI_LOSS_CASE(16, false)
  RLVI_LOSS_CASE(32, true)
  RLVI_LOSS_CASE(32, false)
#undef RLVI_LOSS_CASE
  rlvi_set_error("no loss kernel for fpl=%d", cfg.fpl);
  return RLVI_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int rlvi_loss_f64(rlvi_ctx* ctx, int kind, int intercept, const double* X, const double* y, int64_t n,
                             int d, const double* params, const double* weights, double* losses_out,
                             double* e_out, double* wsum_out, void* stream) {
  RLVI_REQUIRE(ctx && X && params, "null pointer");
  RLVI_REQUIRE(n > 0 && d > 0, "n and d must be positive");
  RLVI_REQUIRE(kind >= RLVI_LOSS_LOGISTIC_CE && kind <= RLVI_LOSS_GAUSSIAN, "unknown loss kind");
  RLVI_REQUIRE(losses_out || e_out || weights, "nothing to compute");
  RLVI_REQUIRE(!weights || wsum_out, "weights given but wsum_out is null");
  if (kind == RLVI_LOSS_LOGISTIC_CE || kind == RLVI_LOSS_SQRES) RLVI_REQUIRE(y != nullptr, "this loss needs y");
  RlviDeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps_per_block = kLossThreads / 32;
  void* scratch = nullptr;

  if (kind == RLVI_LOSS_GAUSSIAN) {
    if (d > 128) {
      rlvi_set_error("RLVI_LOSS_GAUSSIAN supports d <= 128 (got %d)", d);
      return RLVI_ERR_UNSUPPORTED;
    }
    const int nb = (d + 7) / 8, dp = nb * 8;
    const size_t smem = (size_t(dp) * (dp + 8) + dp + 2 * warps_per_block) * sizeof(double);
    const bool vec = rlvi_aligned16(X) && (d % 2 == 0);
    int64_t steps = (n + 7) / 8;
    int64_t want = (steps + warps_per_block - 1) / warps_per_block;
    int grid = int(want < 1 ? 1 : (want > int64_t(ctx->sm_count) * 2 ? int64_t(ctx->sm_count) * 2 : want));
    int rc = rlvi_scratch(ctx, 4096 + size_t(grid) * 2 * sizeof(double), &scratch);
    if (rc != RLVI_OK) return rc;
    GaussParams p;
    p.X = X;
    p.params = params;
    p.w = weights;
    p.losses = losses_out;
    p.e_out = e_out;
    p.wsum_out = wsum_out;
    p.n = n;
    p.d = d;
    p.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + 128);
    p.partials = reinterpret_cast<double*>(static_cast<char*>(scratch) + 4096);
#define RLVI_GAUSS_CASE(NB, V)                                                                             \
  if (nb <= NB && vec == V) {                                                                              \
    RLVI_CUDA(cudaFuncSetAttribute(gaussian_loss_kernel<NB, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   int(smem)));                                                            \
    gaussian_loss_kernel<NB, V><<<grid, kLossThreads, smem, st>>>(p);                                      \
    RLVI_LAUNCH_CHECK(ctx);                                                                                \
    return RLVI_OK;                                                                                        \
  }
    RLVI_GAUSS_CASE(8, true)
    RLVI_GAUSS_CASE(8, false)
    RLVI_GAUSS_CASE(16, true)
    RLVI_GAUSS_CASE(16, false)
#undef RLVI_GAUSS_CASE
    return RLVI_ERR_UNSUPPORTED;
  }

  RowMapCfg cfg;
  if (!rowmap_pick(d, rlvi_aligned16(X), &cfg)) {
    rlvi_set_error("loss kernels support 1 <= d <= 1024 (got %d)", d);
    return RLVI_ERR_UNSUPPORTED;
  }
  const int R = 32 / cfg.L;
  int64_t steps = (n + R - 1) / R;
  int64_t want = (steps + warps_per_block - 1) / warps_per_block;
  int64_t cap = int64_t(ctx->sm_count) * 8;    // 8 resident 256-thread CTAs per SM at <= 32 registers/thread... grid-stride covers the rest
  int grid = int(want < 1 ? 1 : (want > cap ? cap : want));
  int rc = rlvi_scratch(ctx, 4096 + size_t(grid) * 2 * sizeof(double), &scratch);
  if (rc != RLVI_OK) return rc;
  LossParams p;
  p.X = X;
  p.y = y;
  p.params = params;
  p.w = weights;
  p.losses = losses_out;
  p.e_out = e_out;
  p.wsum_out = wsum_out;
  p.n = n;
  p.d = d;
  p.kind = kind;
  p.intercept = intercept ? 1 : 0;
  p.L = cfg.L;
  p.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + 128);
  p.partials = reinterpret_cast<double*>(static_cast<char*>(scratch) + 4096);
  const size_t smem = (size_t(d) + 2 + 2 * warps_per_block) * sizeof(double);
  if (kind == RLVI_LOSS_PCA) return launch_loss<1>(ctx, cfg, p, grid, smem, st);
  if (kind == RLVI_LOSS_SQDIST) return launch_loss<2>(ctx, cfg, p, grid, smem, st);
  return launch_loss<0>(ctx, cfg, p, grid, smem, st);
}
