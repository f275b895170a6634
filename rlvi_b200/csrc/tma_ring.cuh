// tma_ring.cuh -- the producer / private-ring skeleton shared by the TMA-fed streaming kernels
// (loss_tma_kernel, gaussian_tma_kernel in loss.cu; colsum_tma_kernel in moments.cu).
//
// A persistent CTA (one per SM) has up to 16 consumer warps and one producer warp.  The rows of X are cut into
// 32-row tiles; tile t of the CTA (global tile blockIdx.x + t gridDim.x) belongs to consumer warp t % C and is
// its (t / C)-th tile, which lives in that warp's PRIVATE ring of R stages (stage c R + r % R).  Lane c of the
// producer warp feeds consumer c: it waits for the stage to be free and issues ONE cp.async.bulk of the tile's
// 32 d contiguous doubles plus 256 bytes of y and of pi, completing on the stage's mbarrier.
//   * static stage ownership: letting successive uses of one stage go to different warps is unsafe -- a warp could
//     wait for phase k+1 of a barrier still in phase k, which mbarrier.try_wait.parity reports as complete;
//   * one producer lane per consumer: a slow consumer never blocks the refills of the others;
//   * one copy per tile: per-row copies are issue-bound (~60 clk per UBLKCP).
// Stage layout: [32 rows x d doubles][y: 32 doubles][pi: 32 doubles].
#pragma once

#include "tma.cuh"

constexpr int kRingConsumers = 16;
constexpr int kRingThreads = (kRingConsumers + 1) * 32;
constexpr int kRingRows = 32;
constexpr int kRingMaxStages = 2 * kRingConsumers;
constexpr size_t kRingBarrierBytes = 2 * kRingMaxStages * sizeof(uint64_t);

struct RingGeom {
  int stage_bytes;   // kRingRows * d * 8 + 512
  int ncons;         // consumer warps in use (C)
  int depth;         // stages per consumer (R)
};

// Host: fit C * R stages + `tail_bytes` (the kernel's own shared data, barriers included) into 220 KiB.
static inline bool ring_geometry(int d, size_t tail_bytes, RingGeom* g) {
  g->stage_bytes = kRingRows * d * 8 + 512;
  const int max_stages = int((size_t(220) * 1024 - tail_bytes) / size_t(g->stage_bytes));
  g->ncons = max_stages < kRingConsumers ? max_stages : kRingConsumers;
  g->depth = g->ncons > 0 ? max_stages / g->ncons : 0;
  if (g->depth > 2) g->depth = 2;
  return g->ncons >= 2 && g->depth >= 1;
}
__host__ __device__ static inline size_t ring_bytes(const RingGeom& g) { return size_t(g.ncons) * g.depth * g.stage_bytes; }
static inline int ring_grid(const RingGeom& g, int64_t n, int sm_count) {
  const int64_t ntiles = (n + kRingRows - 1) / kRingRows;
  const int64_t want = (ntiles + g.ncons - 1) / g.ncons;
  return int(want < sm_count ? want : sm_count);
}

#ifdef __CUDACC__

struct Ring {
  unsigned char* base;
  uint64_t* full_bar;    // [kRingMaxStages]
  uint64_t* empty_bar;   // [kRingMaxStages]
  RingGeom g;
  int d;
};

// All threads; ends with __syncthreads().
__device__ __forceinline__ void ring_init(const Ring& r) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < r.g.ncons * r.g.depth; ++s) {
      mbar_init(&r.full_bar[s], 1);
      mbar_init(&r.empty_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ int64_t ring_my_tiles(int64_t n) {
  const int64_t ntiles = (n + kRingRows - 1) / kRingRows;
  return (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
}
__device__ __forceinline__ int64_t ring_row0(int64_t t) { return (blockIdx.x + t * gridDim.x) * kRingRows; }

struct RingStage {
  unsigned char* x;   // 32 rows x d doubles, natural pitch
  double* y;          // 32
  double* w;          // 32
  int index;
  uint32_t phase;
};
__device__ __forceinline__ RingStage ring_stage(const Ring& r, int consumer, int64_t t) {
  const int64_t k = t / r.g.ncons;              // k-th tile of this consumer
  RingStage st;
  st.index = consumer * r.g.depth + int(k % r.g.depth);
  st.phase = uint32_t(k / r.g.depth) & 1u;
  st.x = r.base + size_t(st.index) * r.g.stage_bytes;
  st.y = reinterpret_cast<double*>(st.x + size_t(kRingRows) * r.d * 8);
  st.w = st.y + kRingRows;
  return st;
}

// Producer warp (all 32 lanes call; lanes >= C return immediately).  y / w may be null (not copied).
// zero_stale: clear the rows past n of the ragged last tile (needed when they enter a sum with weight 0).
__device__ __forceinline__ void ring_produce(const Ring& r, const double* X, const double* y, const double* w, int64_t n,
                                             int64_t my_tiles, bool zero_stale) {
  const int lane = threadIdx.x & 31;
  if (lane >= r.g.ncons) return;
  const int d = r.d;
  const uint32_t tile_bytes = uint32_t(kRingRows) * uint32_t(d) * 8u;
  for (int64_t t = lane; t < my_tiles; t += r.g.ncons) {
    const RingStage st = ring_stage(r, lane, t);
    mbar_wait(&r.empty_bar[st.index], st.phase ^ 1u);
    const int64_t row0 = ring_row0(t);
    if (row0 + kRingRows <= n) {
      mbar_arrive_expect_tx(&r.full_bar[st.index], tile_bytes + (y ? 256u : 0u) + (w ? 256u : 0u));
      bulk_g2s(st.x, X + row0 * d, tile_bytes, &r.full_bar[st.index]);
      if (y) bulk_g2s(st.y, y + row0, 256, &r.full_bar[st.index]);
      if (w) bulk_g2s(st.w, w + row0, 256, &r.full_bar[st.index]);
    } else {
      // ragged last tile: X rows by one (shorter) bulk copy, y / pi by plain stores that the release semantics
      // of the arrive below publish; rows past n get y = w = 0
      const int rows = int(n - row0);
      for (int i = 0; i < kRingRows; ++i) {
        st.y[i] = (y && i < rows) ? y[row0 + i] : 0.0;
        st.w[i] = (w && i < rows) ? w[row0 + i] : 0.0;
      }
      if (zero_stale) {
        double* xs = reinterpret_cast<double*>(st.x);
        for (int i = rows * d; i < kRingRows * d; ++i) xs[i] = 0.0;
      }
      mbar_arrive_expect_tx(&r.full_bar[st.index], uint32_t(rows) * uint32_t(d) * 8u);
      bulk_g2s(st.x, X + row0 * d, uint32_t(rows) * uint32_t(d) * 8u, &r.full_bar[st.index]);
    }
  }
}

// Consumer side: wait for tile t's stage / hand it back (lane 0 arrives after a __syncwarp()).
__device__ __forceinline__ RingStage ring_acquire(const Ring& r, int consumer, int64_t t) {
  const RingStage st = ring_stage(r, consumer, t);
  mbar_wait(&r.full_bar[st.index], st.phase);
  return st;
}
__device__ __forceinline__ void ring_release(const Ring& r, const RingStage& st) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&r.empty_bar[st.index]);
}

#endif  // __CUDACC__
