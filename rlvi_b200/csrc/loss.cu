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
#include "tma.cuh"

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

// a = x.theta (or ||theta - x||^2 for SQDIST), b = x.x (PCA only), b0 = intercept, yi = label / target.
__device__ __forceinline__ double finish_loss(int kind, double a, double b, double b0, double yi) {
  switch (kind) {
    case RLVI_LOSS_LOGISTIC_CE: {
      const double phi = b0 + a;
      return (-yi * phi + phi) + log1p(exp(-phi));          // utils.py:21, same operation order
    }
    case RLVI_LOSS_SOFTPLUS: {
      const double phi = b0 + a;
      return fmax(phi, 0.0) + log1p(exp(-fabs(phi)));        // logaddexp(0, phi)
    }
    case RLVI_LOSS_SQRES: {
      const double r = yi - (b0 + a);
      return r * r;
    }
    case RLVI_LOSS_SQDIST: {
      const double r = sqrt(a);                              // np.linalg.norm(...)**2
      return r * r;
    }
    default:                                                 // RLVI_LOSS_PCA
      return b - a * a;
  }
}

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
      const double loss = finish_loss(p.kind, a, b, b0, (p.kind == RLVI_LOSS_LOGISTIC_CE || p.kind == RLVI_LOSS_SQRES) ? p.y[row] : 0.0);
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
// TMA-fed streaming kernel (d % 16 == 0, d <= 256, 16-byte aligned pointers): the headline path.
//   * persistent, one CTA per SM: 16 consumer warps + 1 producer warp;
//   * the producer streams 32-row tiles into an S-stage shared-memory ring: ONE cp.async.bulk (TMA engine,
//     SASS UBLKCP) of 32*d*8 contiguous bytes of X per tile, plus 256 bytes of y and of pi, completing on
//     an mbarrier.  (A first version issued one 512-byte copy per row into a padded pitch: the serialised
//     per-row UBLKCP issue, ~67 clk per row, capped the kernel at 2.3 TB/s.)
//   * tile t of the CTA is consumed by warp t mod C (C <= 16 consumer warps, each with a private ring of
//     R stages; C*R stages fill the shared memory): lane l owns row l of the tile and does the whole
//     d-long dot product itself.  Rows sit at their natural d*8-byte pitch (a multiple of 128 bytes), so
//     lane l walks its row ROTATED by l 16-byte units -- unit (j + l) mod (d/2) at step j -- which makes
//     every 128-bit shared-memory read of x and of theta bank-conflict free;
//   * the exp / log1p tail then runs with ALL 32 lanes busy (the register-tiled kernel above leaves 3/4
//     of the lanes idle there and exposes the global-load latency once per row group);
//   * outputs are one coalesced 256-byte store per warp and array.
// Algorithmic bytes per sample: d*8 (X) + 8 (y) [+ 8 (pi)] in, 8 (l) and/or 8 (e) out.
// ---------------------------------------------------------------------------------------------
constexpr int kTmaConsumers = 16;
constexpr int kTmaThreads = (kTmaConsumers + 1) * 32;
constexpr int kTmaRows = 32;
constexpr int kTmaMaxStages = 2 * kTmaConsumers;

struct LossTmaParams {
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
  int stage_bytes;    // kTmaRows * d * 8 + 512 (y, pi)
  int ncons;          // consumer warps in use (<= kTmaConsumers)
  int depth;          // ring stages per consumer warp; nstages = ncons * depth
  int nstages;
  double* partials;   // [grid][2]
  unsigned int* ticket;
};

template <int DOTK>
__global__ void __launch_bounds__(kTmaThreads, 1) loss_tma_kernel(const LossTmaParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = p.nstages, C = p.ncons, R = p.depth;
  unsigned char* ring = smem_raw;
  double* sTheta = reinterpret_cast<double*>(ring + size_t(S) * p.stage_bytes);      // d (+ pad to even)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sTheta + ((p.d + 1) & ~1) + 2);
  uint64_t* empty_bar = full_bar + kTmaMaxStages;
  double* sRed = reinterpret_cast<double*>(empty_bar + kTmaMaxStages);               // 2 * 17 doubles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = p.d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) sTheta[i] = p.params[i + (p.intercept ? 1 : 0)];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double b0 = p.intercept ? p.params[0] : 0.0;
  const int64_t ntiles = (p.n + kTmaRows - 1) / kTmaRows;
  // tiles of this CTA: blockIdx.x, blockIdx.x + grid, ...   local index t -> stage t % S, warp t % 8
  const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const bool need_y = (p.kind == RLVI_LOSS_LOGISTIC_CE || p.kind == RLVI_LOSS_SQRES);
  double s_wl = 0.0, s_w = 0.0;

  if (warp == kTmaConsumers) {
    // ===== producer warp =========================================================================
    // Lane c of the producer warp feeds consumer warp c (its private ring of R stages), so a slow consumer
    // never blocks the refills of the others.  Stage ownership is static: the r-th tile of consumer c is
    // local tile c + r C and lives in stage c R + r % R.  (Letting successive uses of one stage go to
    // different warps is unsafe: a warp could wait for phase k+1 of a barrier still in phase k, which
    // mbarrier.try_wait.parity reports as complete.)
    const uint32_t tile_bytes = uint32_t(kTmaRows) * uint32_t(d) * 8u;
    if (lane < C) {
      int64_t r = 0;
      for (int64_t t = lane; t < my_tiles; t += C, ++r) {
        const int stage = lane * R + int(r % R);
        const uint32_t phase = uint32_t(r / R) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        unsigned char* sX = ring + size_t(stage) * p.stage_bytes;
        double* sY = reinterpret_cast<double*>(sX + tile_bytes);
        double* sW = sY + kTmaRows;
        const int64_t row0 = (blockIdx.x + t * gridDim.x) * kTmaRows;
        if (row0 + kTmaRows <= p.n) {
          mbar_arrive_expect_tx(&full_bar[stage], tile_bytes + (need_y ? 256u : 0u) + (p.w ? 256u : 0u));
          bulk_g2s(sX, p.X + row0 * d, tile_bytes, &full_bar[stage]);
          if (need_y) bulk_g2s(sY, p.y + row0, 256, &full_bar[stage]);
          if (p.w) bulk_g2s(sW, p.w + row0, 256, &full_bar[stage]);
        } else {
          // ragged last tile: X rows by one (shorter) bulk copy, y / pi by plain stores that the
          // release semantics of the arrive below publish (rows past n are never read back)
          const int rows = int(p.n - row0);
          for (int i = 0; i < rows; ++i) {
            sY[i] = need_y ? p.y[row0 + i] : 0.0;
            sW[i] = p.w ? p.w[row0 + i] : 0.0;
          }
          mbar_arrive_expect_tx(&full_bar[stage], uint32_t(rows) * uint32_t(d) * 8u);
          bulk_g2s(sX, p.X + row0 * d, uint32_t(rows) * uint32_t(d) * 8u, &full_bar[stage]);
        }
      }
    }
  } else if (warp < C) {
    // ===== consumer warps ========================================================================
    const double2* th2 = reinterpret_cast<const double2*>(sTheta);
    const int npair = d >> 1;                      // 16-byte units per row, a multiple of 8
    const uint32_t tile_bytes = uint32_t(kTmaRows) * uint32_t(d) * 8u;
    for (int64_t t = warp; t < my_tiles; t += C) {
      const int64_t r = t / C;
      const int stage = warp * R + int(r % R);
      const uint32_t phase = uint32_t(r / R) & 1u;
      mbar_wait(&full_bar[stage], phase);
      const unsigned char* sX = ring + size_t(stage) * p.stage_bytes;
      const double* sY = reinterpret_cast<const double*>(sX + tile_bytes);
      const double* sW = sY + kTmaRows;
      const double2* xr = reinterpret_cast<const double2*>(sX) + size_t(lane) * npair;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, q0 = 0.0, q1 = 0.0;
      int u = lane % npair;                         // rotated walk: conflict-free at the natural pitch
#pragma unroll 4
      for (int j = 0; j < npair; j += 2) {
        int u1 = u + 1;
        if (u1 == npair) u1 = 0;
        const double2 x0 = xr[u], x1 = xr[u1];
        const double2 t0 = th2[u], t1 = th2[u1];
        u = u1 + 1;
        if (u == npair) u = 0;
        if (DOTK == 2) {
          const double v0 = t0.x - x0.x, v1 = t0.y - x0.y, v2 = t1.x - x1.x, v3 = t1.y - x1.y;
          a0 = fma(v0, v0, a0);
          a1 = fma(v1, v1, a1);
          a2 = fma(v2, v2, a2);
          a3 = fma(v3, v3, a3);
        } else {
          a0 = fma(x0.x, t0.x, a0);
          a1 = fma(x0.y, t0.y, a1);
          a2 = fma(x1.x, t1.x, a2);
          a3 = fma(x1.y, t1.y, a3);
          if (DOTK == 1) {
            q0 = fma(x0.x, x0.x, q0);
            q1 = fma(x0.y, x0.y, q1);
            q0 = fma(x1.x, x1.x, q0);
            q1 = fma(x1.y, x1.y, q1);
          }
        }
      }
      const double a = (a0 + a1) + (a2 + a3);
      const double b = q0 + q1;
      const double yi = sY[lane];
      const double wi = sW[lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);      // the stage can be refilled while we do the tail
      const int64_t row = (blockIdx.x + t * gridDim.x) * kTmaRows + lane;
      if (row < p.n) {
        const double loss = finish_loss(p.kind, a, b, b0, need_y ? yi : 0.0);
        if (p.losses) p.losses[row] = loss;
        if (p.e_out) p.e_out[row] = exp(-loss);
        if (p.w) {
          s_wl = fma(wi, loss, s_wl);
          s_w += wi;
        }
      }
    }
  }

  if (p.w) {   // uniform across the grid
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

// ---------------------------------------------------------------------------------------------
// TMA-fed Gaussian NLL (d % 16 == 0, d <= 64, aligned): same producer / private-ring skeleton as
// loss_tma_kernel; a consumer warp walks its 32-row tile as four 8-row DMMA groups, gathers the four groups'
// quadratic forms so that lane l ends with row l (one shuffle per group), and then runs the exp tail and the
// 256-byte stores with all 32 lanes.  (gaussian_loss_kernel above fetches its fragments straight from global
// memory: latency-bound at 47 % of the DMMA peak.)
// ---------------------------------------------------------------------------------------------
struct GaussTmaParams {
  const double* X;
  const double* params;   // [c, mu(d), U(d*d) row-major upper triangular]
  const double* w;
  double* losses;
  double* e_out;
  double* wsum_out;
  int64_t n;
  int d;
  int stage_bytes;
  int ncons, depth;
  double* partials;
  unsigned int* ticket;
};

template <int NB>   // NB = d / 8 MMA blocks (exact)
__global__ void __launch_bounds__(kTmaThreads, 1) gaussian_tma_kernel(const GaussTmaParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int D = NB * 8;
  constexpr int kPitchU = D + 8;                 // doubles: conflict-free 128-bit reads of U rows
  const int C = p.ncons, R = p.depth, S = C * R;
  unsigned char* ring = smem_raw;
  double* sU = reinterpret_cast<double*>(ring + size_t(S) * p.stage_bytes);   // D * kPitchU
  double* sMu = sU + D * kPitchU;                                              // D
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sMu + D);
  uint64_t* empty_bar = full_bar + kTmaMaxStages;
  double* sRed = reinterpret_cast<double*>(empty_bar + kTmaMaxStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < D * kPitchU; i += blockDim.x) {
    const int r = i / kPitchU, c = i - r * kPitchU;
    sU[i] = (c < D && c >= r) ? p.params[1 + D + size_t(r) * D + c] : 0.0;
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) sMu[i] = p.params[1 + i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double cst = p.params[0];
  const int64_t ntiles = (p.n + kTmaRows - 1) / kTmaRows;
  const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t tile_bytes = uint32_t(kTmaRows) * uint32_t(D) * 8u;
  double s_wl = 0.0, s_w = 0.0;

  if (warp == kTmaConsumers) {
    if (lane < C) {
      int64_t r = 0;
      for (int64_t t = lane; t < my_tiles; t += C, ++r) {
        const int stage = lane * R + int(r % R);
        const uint32_t phase = uint32_t(r / R) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        unsigned char* sX = ring + size_t(stage) * p.stage_bytes;
        double* sW = reinterpret_cast<double*>(sX + tile_bytes);
        const int64_t row0 = (blockIdx.x + t * gridDim.x) * kTmaRows;
        if (row0 + kTmaRows <= p.n) {
          mbar_arrive_expect_tx(&full_bar[stage], tile_bytes + (p.w ? 256u : 0u));
          bulk_g2s(sX, p.X + row0 * D, tile_bytes, &full_bar[stage]);
          if (p.w) bulk_g2s(sW, p.w + row0, 256, &full_bar[stage]);
        } else {
          const int rows = int(p.n - row0);
          for (int i = 0; i < rows; ++i) sW[i] = p.w ? p.w[row0 + i] : 0.0;
          mbar_arrive_expect_tx(&full_bar[stage], uint32_t(rows) * uint32_t(D) * 8u);
          bulk_g2s(sX, p.X + row0 * D, uint32_t(rows) * uint32_t(D) * 8u, &full_bar[stage]);
        }
      }
    }
  } else if (warp < C) {
    const int g = lane >> 2, t = lane & 3;
    for (int64_t tl = warp; tl < my_tiles; tl += C) {
      const int64_t r = tl / C;
      const int stage = warp * R + int(r % R);
      const uint32_t phase = uint32_t(r / R) & 1u;
      mbar_wait(&full_bar[stage], phase);
      const unsigned char* sX = ring + size_t(stage) * p.stage_bytes;
      const double* sW = reinterpret_cast<const double*>(sX + tile_bytes);
      double myquad = 0.0;
#pragma unroll 1
      for (int grp = 0; grp < 4; ++grp) {
        // centred fragment of row 8 grp + g: features (8 kk + 2 t, + 1)
        const double2* xrow = reinterpret_cast<const double2*>(sX + size_t(8 * grp + g) * D * 8);
        double2 c[NB];
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) {
          double2 v = xrow[4 * kk + t];
          const double2 m = *reinterpret_cast<const double2*>(sMu + 8 * kk + 2 * t);
          v.x -= m.x;
          v.y -= m.y;
          c[kk] = v;
        }
        double quad = 0.0;
#pragma unroll
        for (int nbk = 0; nbk < NB; ++nbk) {
          double z0 = 0.0, z1 = 0.0;
          const double* urow = sU + size_t(8 * nbk + g) * kPitchU + 2 * t;
#pragma unroll
          for (int kk = nbk; kk < NB; ++kk) {
            const double2 u = *reinterpret_cast<const double2*>(urow + 8 * kk);
            dmma884(z0, z1, c[kk].x, u.x);
            dmma884(z0, z1, c[kk].y, u.y);
          }
          quad = fma(z0, z0, quad);
          quad = fma(z1, z1, quad);
        }
        quad += __shfl_xor_sync(0xffffffffu, quad, 1);
        quad += __shfl_xor_sync(0xffffffffu, quad, 2);
        // row 8 grp + g's form sits in lanes 4 g .. 4 g + 3; lane l wants row l
        const double got = __shfl_sync(0xffffffffu, quad, 4 * (lane & 7));
        if ((lane >> 3) == grp) myquad = got;
      }
      const double wi = sW[lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      const int64_t row = (blockIdx.x + tl * gridDim.x) * kTmaRows + lane;
      if (row < p.n) {
        const double loss = 0.5 * (myquad + cst);                    // utils.py:101
        if (p.losses) p.losses[row] = loss;
        if (p.e_out) p.e_out[row] = exp(-loss);
        if (p.w) {
          s_wl = fma(wi, loss, s_wl);
          s_w += wi;
        }
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
    // ---- TMA-fed path: d in {16, 32, 48, 64}, aligned, at least one full tile ------------------------
    if (d % 16 == 0 && d <= 64 && n >= kTmaRows && rlvi_aligned16(X) && (!weights || rlvi_aligned16(weights))) {
      GaussTmaParams q;
      q.X = X;
      q.params = params;
      q.w = weights;
      q.losses = losses_out;
      q.e_out = e_out;
      q.wsum_out = wsum_out;
      q.n = n;
      q.d = d;
      q.stage_bytes = kTmaRows * d * 8 + 256;
      const size_t tail = (size_t(d) * (d + 8) + d) * 8 + 2 * kTmaMaxStages * 8 + 2 * (kTmaConsumers + 1) * 8 + 128;
      const int max_stages = int((size_t(220) * 1024 - tail) / q.stage_bytes);
      q.ncons = max_stages < kTmaConsumers ? max_stages : kTmaConsumers;
      q.depth = q.ncons > 0 ? max_stages / q.ncons : 0;
      if (q.depth > 2) q.depth = 2;
      const int64_t ntiles = (n + kTmaRows - 1) / kTmaRows;
      const int64_t want_ctas = (ntiles + q.ncons - 1) / q.ncons;
      const int grid = int(want_ctas < ctx->sm_count ? want_ctas : ctx->sm_count);
      int rc = rlvi_scratch(ctx, 4096 + size_t(grid) * 2 * sizeof(double), &scratch);
      if (rc != RLVI_OK) return rc;
      q.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + 128);
      q.partials = reinterpret_cast<double*>(static_cast<char*>(scratch) + 4096);
      const size_t smem = size_t(q.ncons) * q.depth * q.stage_bytes + tail;
#define RLVI_GTMA_CASE(NB)                                                                                    \
  if (d == 8 * NB) {                                                                                          \
    RLVI_CUDA(cudaFuncSetAttribute(gaussian_tma_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
    gaussian_tma_kernel<NB><<<grid, kTmaThreads, smem, st>>>(q);                                              \
    RLVI_LAUNCH_CHECK(ctx);                                                                                   \
    return RLVI_OK;                                                                                           \
  }
      RLVI_GTMA_CASE(2)
      RLVI_GTMA_CASE(4)
      RLVI_GTMA_CASE(6)
      RLVI_GTMA_CASE(8)
#undef RLVI_GTMA_CASE
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

  // ---- TMA-fed path: d % 16 == 0, <= 256, all streamed arrays 16-byte aligned, at least one full tile ----
  if (d % 16 == 0 && d <= 256 && n >= kTmaRows && rlvi_aligned16(X) && (!y || rlvi_aligned16(y)) &&
      (!weights || rlvi_aligned16(weights))) {
    LossTmaParams q;
    q.X = X;
    q.y = y;
    q.params = params;
    q.w = weights;
    q.losses = losses_out;
    q.e_out = e_out;
    q.wsum_out = wsum_out;
    q.n = n;
    q.d = d;
    q.kind = kind;
    q.intercept = intercept ? 1 : 0;
    q.stage_bytes = kTmaRows * d * 8 + 512;
    const size_t tail = size_t(((d + 1) & ~1) + 2) * 8 + 2 * kTmaMaxStages * 8 + 2 * (kTmaConsumers + 1) * 8 + 128;
    const int max_stages = int((size_t(220) * 1024 - tail) / q.stage_bytes);
    q.ncons = max_stages < kTmaConsumers ? max_stages : kTmaConsumers;
    q.depth = q.ncons > 0 ? max_stages / q.ncons : 0;
    if (q.depth > 2) q.depth = 2;
    const int stages = q.ncons * q.depth;
    const int64_t ntiles = (n + kTmaRows - 1) / kTmaRows;
    if (stages >= 2) {
      q.nstages = stages;
      const int64_t want_ctas = (ntiles + q.ncons - 1) / q.ncons;
      const int grid = int(want_ctas < ctx->sm_count ? want_ctas : ctx->sm_count);
      int rc = rlvi_scratch(ctx, 4096 + size_t(grid) * 2 * sizeof(double), &scratch);
      if (rc != RLVI_OK) return rc;
      q.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + 128);
      q.partials = reinterpret_cast<double*>(static_cast<char*>(scratch) + 4096);
      const size_t smem = size_t(stages) * q.stage_bytes + tail;
#define RLVI_TMA_CASE(K)                                                                                   \
  {                                                                                                        \
    RLVI_CUDA(cudaFuncSetAttribute(loss_tma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
    loss_tma_kernel<K><<<grid, kTmaThreads, smem, st>>>(q);                                                \
    RLVI_LAUNCH_CHECK(ctx);                                                                                \
    return RLVI_OK;                                                                                        \
  }
      if (kind == RLVI_LOSS_PCA) RLVI_TMA_CASE(1)
      if (kind == RLVI_LOSS_SQDIST) RLVI_TMA_CASE(2)
      RLVI_TMA_CASE(0)
#undef RLVI_TMA_CASE
    }
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
