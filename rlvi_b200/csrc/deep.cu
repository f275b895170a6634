// deep.cu -- the deep-learning RLVI pieces in FP32.
//
//   rlvi_wce_fwd_bwd_f32   deep-learning/methods/train_rlvi.py:89-94 + the autograd backward of line 96
//                          (and utils.py:65-79 `accuracy`, called at train_rlvi.py:85) in ONE launch
//                          instead of ~12 ATen kernels.
//   rlvi_fn_threshold_f32  train_rlvi.py:41-49 (false_negative_criterion) + :102-103 (truncation)
//                          as one single-CTA radix select instead of sort + cumsum + compare + index.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kWceThreads = 256;

struct WceParams {
  const float* logits;
  const int64_t* labels;
  const int64_t* indexes;
  const float* weights;
  float* residuals;
  float* per_sample;
  float* dlogits;
  float* out_loss;
  int32_t* out_correct;
  int64_t batch;
  int classes;
  int64_t n_train;
  double* partials;
  unsigned int* ticket;
};

// One warp per row; NV logits per lane held in registers (classes <= 32 * NV).
template <int NV>
__global__ void __launch_bounds__(kWceThreads) wce_kernel(const WceParams p) {
  __shared__ double s_red[kWceThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int C = p.classes;
  const float inv_b = 1.0f / float(p.batch);
  double wl_sum = 0.0;   // sum_i loss_i * w_i over this warp's rows (lane 0)
  int top1 = 0, top5 = 0;
  for (int64_t row = int64_t(blockIdx.x) * nwarp + warp; row < p.batch; row += int64_t(gridDim.x) * nwarp) {
    const float* x = p.logits + row * C;
    float v[NV];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      v[k] = (c < C) ? __ldg(x + c) : -INFINITY;
      m = fmaxf(m, v[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      if (c < C) s += expf(v[k] - m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float logs = logf(s);
    const int64_t label = p.labels[row];
    const int64_t idx = p.indexes ? p.indexes[row] : row;
    const float xl = __ldg(x + label);
    const float loss = -((xl - m) - logs);                 // -log_softmax[label]
    const float w = (idx >= 0 && idx < p.n_train) ? p.weights[idx] : 0.f;
    if (lane == 0) {
      if (idx >= 0 && idx < p.n_train) p.residuals[idx] = loss;   // line 90, detached (quirk Q8)
      if (p.per_sample) p.per_sample[row] = loss;
      wl_sum += double(loss * w);                                  // line 93: loss * batch_weights in FP32
    }
    if (p.dlogits) {
      const float gsc = w * inv_b;                                 // d mean / d loss_i, times pi_i
      float* dx = p.dlogits + row * C;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < C) {
          float d = expf((v[k] - m) - logs) * gsc;
          if (c == label) d -= gsc;
          dx[c] = d;
        }
      }
    }
    if (p.out_correct) {
      // rank of the label among the logits (ties: lower class index first)
      int rank = 0;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < C && (v[k] > xl || (v[k] == xl && c < label))) ++rank;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
      if (lane == 0) {
        top1 += (rank < 1);
        top5 += (rank < 5);
      }
    }
  }
  // block partial of the weighted loss: warps in order
  if (lane == 0) s_red[warp] = wl_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < nwarp; ++w) s += s_red[w];
    p.partials[blockIdx.x] = s;
  }
  if (p.out_correct && lane == 0 && (top1 | top5)) {
    atomicAdd(p.out_correct, top1);        // integer atomics: order-independent
    atomicAdd(p.out_correct + 1, top5);
  }
  if (last_block_ticket(p.ticket, gridDim.x)) {
    if (threadIdx.x < 32) {
      double a = 0.0;
      for (unsigned int j = lane; j < gridDim.x; j += 32) a += p.partials[j];
      a = warp_sum(a);
      if (lane == 0) p.out_loss[0] = float(a / double(p.batch));   // .mean() (line 94)
    }
  }
}

// ---------------------------------------------------------------------------------------------
// false-negative threshold: radix select over the FP32 bit patterns, masses in 2^-40 fixed point
// ---------------------------------------------------------------------------------------------
constexpr int kThrThreads = 1024;
constexpr double kFix = 1099511627776.0;   // 2^40

struct ThrShared {
  unsigned int cnt[256];
  unsigned long long mass[256];
  unsigned long long red_mass[kThrThreads / 32];
  unsigned int red_key[kThrThreads / 32];
  // scalars broadcast by thread 0
  unsigned int prefix;
  unsigned long long P, M;
  int done;
  unsigned int result_key;
  int need_pred;     // answer is the smallest key > prefix
  int need_min;      // answer is the smallest weight
};

__device__ __forceinline__ unsigned long long mass_of(float w) {
  const float u = 1.0f - w;                      // the FP32 value torch forms (line 46)
  return u > 0.f ? (unsigned long long)__double2ll_rn(double(u) * kFix) : 0ull;
}

__global__ void __launch_bounds__(kThrThreads) fn_threshold_kernel(float* weights, int64_t n, float alpha,
                                                                   float prev_threshold, int truncate,
                                                                   float* out_threshold) {
  __shared__ ThrShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- pass 0: total mass and the smallest key
  unsigned long long tm = 0ull;
  unsigned int kmin = 0xffffffffu;
  for (int64_t i = tid; i < n; i += kThrThreads) {
    const float w = weights[i];
    tm += mass_of(w);
    kmin = min(kmin, __float_as_uint(w));
  }
  for (int o = 16; o > 0; o >>= 1) {
    tm += __shfl_xor_sync(0xffffffffu, tm, o);
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
  }
  if (lane == 0) {
    sh.red_mass[warp] = tm;
    sh.red_key[warp] = kmin;
  }
  __syncthreads();
  unsigned long long beta_fix = 0ull;
  unsigned int key_min = 0xffffffffu;
  {
    unsigned long long t = 0ull;
    for (int w = 0; w < kThrThreads / 32; ++w) {
      t += sh.red_mass[w];
      key_min = min(key_min, sh.red_key[w]);
    }
    const float total = float(double(t) / kFix);          // torch.sum(1 - weights): FP32 scalar
    const float beta = total * alpha;                     // line 44
    beta_fix = (unsigned long long)(double(beta) * kFix);
  }
  if (tid == 0) {
    sh.prefix = 0u;
    sh.P = 0ull;
    sh.M = 0ull;
    sh.done = 0;
    sh.need_pred = 0;
    sh.need_min = 0;
    sh.result_key = 0u;
  }
  __syncthreads();

  // ---- four 8-bit levels, most significant first (weights >= 0: bit pattern order == value order)
  for (int level = 0; level < 4 && !sh.done; ++level) {
    const int shift = 24 - 8 * level;
    const unsigned int himask = level == 0 ? 0u : (0xffffffffu << (shift + 8));
    if (tid < 256) {
      sh.cnt[tid] = 0u;
      sh.mass[tid] = 0ull;
    }
    __syncthreads();
    const unsigned int prefix = sh.prefix;
    for (int64_t i = tid; i < n; i += kThrThreads) {
      const float w = weights[i];
      const unsigned int key = __float_as_uint(w);
      if ((key & himask) == prefix) {
        const unsigned int b = (key >> shift) & 255u;
        atomicAdd(&sh.cnt[b], 1u);
        atomicAdd(&sh.mass[b], mass_of(w));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long P = sh.P, M = sh.M;
      int b = 255;
      for (; b >= 0; --b) {
        if (M + sh.mass[b] > beta_fix) break;   // this bucket does not fit entirely
        P += sh.cnt[b];
        M += sh.mass[b];
      }
      sh.P = P;
      sh.M = M;
      if (b < 0) {
        // everything under this prefix fits: the boundary is at the end of the prefix range
        if (level == 0) {
          sh.need_min = 1;           // all n weights fit: index n-1 -> the smallest weight
        } else {
          // cannot happen: the parent level chose this bucket because it did NOT fit entirely
          sh.need_min = 1;
        }
        sh.done = 1;
      } else {
        sh.prefix = prefix | (unsigned int)(b) << shift;
        if (level == 3) {
          const unsigned int key = sh.prefix;
          const unsigned long long u = mass_of(__uint_as_float(key));
          const unsigned long long room = beta_fix - M;          // M <= beta_fix here
          unsigned long long t = (u == 0ull) ? sh.cnt[b] : room / u;
          if (t > sh.cnt[b]) t = sh.cnt[b];
          if (t >= 1ull) {
            sh.result_key = key;
          } else if (P == 0ull) {
            sh.need_min = 1;          // nothing fits: index -1 wraps to the smallest weight (Q9)
          } else {
            sh.need_pred = 1;         // boundary falls just before this key: previous (larger) key
          }
          sh.done = 1;
        }
      }
    }
    __syncthreads();
  }

  if (sh.need_pred) {
    const unsigned int key = sh.prefix;
    unsigned int best = 0xffffffffu;
    for (int64_t i = tid; i < n; i += kThrThreads) {
      const unsigned int k = __float_as_uint(weights[i]);
      if (k > key) best = min(best, k);
    }
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    __syncthreads();
    if (lane == 0) sh.red_key[warp] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned int b = 0xffffffffu;
      for (int w = 0; w < kThrThreads / 32; ++w) b = min(b, sh.red_key[w]);
      sh.result_key = b;
    }
    __syncthreads();
  } else if (sh.need_min) {
    if (tid == 0) sh.result_key = key_min;
    __syncthreads();
  }
  const float thr_new = __uint_as_float(sh.result_key);
  const float thr = fmaxf(prev_threshold, thr_new);          // line 102
  if (tid == 0) out_threshold[0] = thr;
  if (truncate) {
    for (int64_t i = tid; i < n; i += kThrThreads) {
      if (weights[i] < thr) weights[i] = 0.f;                // line 103
    }
  }
}

// ---------------------------------------------------------------------------------------------
// KKT shift objective (standard-learning/rlvi.py:34-42)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) shift_sum_kernel(const double* __restrict__ losses, int64_t n, double shift,
                                                        double c, double* pi_out, double* partials,
                                                        unsigned int* ticket, double* out_sum) {
  __shared__ double s_red[8];
  double acc = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double t = exp(-losses[i] + shift);
    const double r = t / (c + t);
    acc += r;
    if (pi_out) pi_out[i] = r;
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

}  // namespace

extern "C" int rlvi_wce_fwd_bwd_f32(rlvi_ctx* ctx, const float* logits, const int64_t* labels,
                                    const int64_t* indexes, const float* weights, float* residuals, int64_t batch,
                                    int classes, int64_t n_train, float* per_sample_out, float* dlogits,
                                    float* out_loss, int32_t* out_correct, void* stream) {
  RLVI_REQUIRE(ctx && logits && labels && weights && residuals && out_loss, "null pointer");
  RLVI_REQUIRE(batch > 0 && classes > 0 && n_train > 0, "batch, classes, n_train must be positive");
  if (classes > 1024) {
    rlvi_set_error("rlvi_wce_fwd_bwd_f32 supports classes <= 1024 (got %d)", classes);
    return RLVI_ERR_UNSUPPORTED;
  }
  RlviDeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nwarp = kWceThreads / 32;
  int64_t want = (batch + nwarp - 1) / nwarp;
  const int64_t cap = int64_t(ctx->sm_count) * 8;
  const int grid = int(want > cap ? cap : want);
  void* scratch = nullptr;
  int rc = rlvi_scratch(ctx, 4096 + size_t(grid) * sizeof(double), &scratch);
  if (rc != RLVI_OK) return rc;
  WceParams p;
  p.logits = logits;
  p.labels = labels;
  p.indexes = indexes;
  p.weights = weights;
  p.residuals = residuals;
  p.per_sample = per_sample_out;
  p.dlogits = dlogits;
  p.out_loss = out_loss;
  p.out_correct = out_correct;
  p.batch = batch;
  p.classes = classes;
  p.n_train = n_train;
  p.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + 128);
  p.partials = reinterpret_cast<double*>(static_cast<char*>(scratch) + 4096);
  if (out_correct) RLVI_CUDA(cudaMemsetAsync(out_correct, 0, 2 * sizeof(int32_t), st));
  if (classes <= 128)
    wce_kernel<4><<<grid, kWceThreads, 0, st>>>(p);
  else if (classes <= 256)
    wce_kernel<8><<<grid, kWceThreads, 0, st>>>(p);
  else
    wce_kernel<32><<<grid, kWceThreads, 0, st>>>(p);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}

extern "C" int rlvi_fn_threshold_f32(rlvi_ctx* ctx, float* weights, int64_t n, float alpha, float prev_threshold,
                                     int truncate, float* out_threshold, void* stream) {
  RLVI_REQUIRE(ctx && weights && out_threshold, "null pointer");
  RLVI_REQUIRE(n > 0, "n must be positive");
  RLVI_REQUIRE(n < (int64_t(1) << 22), "single-CTA selection supports n < 2^22");
  RlviDeviceGuard guard(ctx->device);
  fn_threshold_kernel<<<1, kThrThreads, 0, static_cast<cudaStream_t>(stream)>>>(weights, n, alpha, prev_threshold,
                                                                              truncate, out_threshold);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}

extern "C" int rlvi_shift_sum_f64(rlvi_ctx* ctx, const double* losses, int64_t n, double shift, double c,
                                  double* pi_out, double* out_sum, void* stream) {
  RLVI_REQUIRE(ctx && losses && out_sum, "null pointer");
  RLVI_REQUIRE(n > 0, "n must be positive");
  RlviDeviceGuard guard(ctx->device);
  int64_t want = (n + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = int64_t(ctx->sm_count) * 8;
  const int grid = int(want < 1 ? 1 : (want > cap ? cap : want));
  void* scratch = nullptr;
  int rc = rlvi_scratch(ctx, 4096 + size_t(grid) * sizeof(double), &scratch);
  if (rc != RLVI_OK) return rc;
  shift_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      losses, n, shift, c, pi_out, reinterpret_cast<double*>(static_cast<char*>(scratch) + 4096),
      reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + 128), out_sum);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}
