// rowmap.cuh -- how a warp walks the rows of a row-major FP64 matrix X[n][d].
//
// L lanes (a power of two, 1..32) share one row; a warp covers 32/L rows per step.  Lane q of a row
// group owns FPL feature slots.  With VEC (d even, rows 16-byte aligned) slot k is feature
// 2*(q + L*(k/2)) + (k&1): the L lanes of a group issue one contiguous L*16-byte 128-bit request per
// pair, so every 32-byte sector fetched is fully used.  Without VEC slot k is feature q + L*k.
// Slots beyond d (and rows beyond n) read as zero.
#pragma once

#include "common.cuh"

struct RowMapCfg {
  int L;      // lanes per row
  int fpl;    // feature slots per lane: 4, 16 or 32
  bool vec;
};

// d <= 32 -> 4 slots; d <= 512 -> 16 slots; d <= 1024 -> 32 slots.
static inline bool rowmap_pick(int d, bool aligned, RowMapCfg* cfg) {
  if (d < 1 || d > 1024) return false;
  int fpl = d <= 32 ? 4 : (d <= 512 ? 16 : 32);
  int L = 1;
  while (L * fpl < d) L <<= 1;
  cfg->L = L;
  cfg->fpl = fpl;
  cfg->vec = aligned && (d % 2 == 0);
  return true;
}

#ifdef __CUDACC__

template <int FPL, bool VEC>
struct RowMap {
  // feature index of slot k for lane-in-group q
  __device__ __forceinline__ static int feature(int k, int q, int L) {
    return VEC ? (2 * (q + L * (k >> 1)) + (k & 1)) : (q + L * k);
  }
  // load this lane's slots of one row (zeros where the slot is past d or the row is invalid)
  __device__ __forceinline__ static void load_row(const double* __restrict__ X, int64_t row, int d, int q, int L,
                                                  bool valid, double (&x)[FPL]) {
    const double* base = X + row * int64_t(d);
    if (VEC) {
#pragma unroll
      for (int k = 0; k < FPL; k += 2) {
        const int f = 2 * (q + L * (k >> 1));
        if (valid && f < d) {
          const double2 v = ld_stream_d2(base + f);
          x[k] = v.x;
          x[k + 1] = v.y;
        } else {
          x[k] = 0.0;
          x[k + 1] = 0.0;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < FPL; ++k) {
        const int f = q + L * k;
        x[k] = (valid && f < d) ? ld_stream_d1(base + f) : 0.0;
      }
    }
  }
  // load a length-d vector (e.g. theta) from shared memory into the same slots
  __device__ __forceinline__ static void load_vec(const double* v, int d, int q, int L, double (&t)[FPL]) {
#pragma unroll
    for (int k = 0; k < FPL; ++k) {
      const int f = feature(k, q, L);
      t[k] = (f < d) ? v[f] : 0.0;
    }
  }
};

// sum over the L lanes of a row group (butterfly: every lane of the group ends with the total)
__device__ __forceinline__ double group_sum(double v, int L) {
  for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__
