// gram_tf32_pair.cu -- the TF32 weighted Gram of FP32-stored samples on CTA PAIRS (tcgen05.mma.cta_group::2, M = 256):
// the d = 129..256 and d = 385..512 shapes of rlvi_weighted_moments_f32 (config C3: N = 2^24, d = 512).  Same
// statistics, same three-level accumulation and the same operand layout as gram_tf32.cu (read its header first);
// what changes is who computes what:
//
//   standard-learning/utils.py:82-84   PCA of the rows pi_i x_i      (power = 2:  G = sum pi_i^2 x_i x_i^T)
//   standard-learning/rlvi.py:70-71    sqrt(pi)-scaled least squares (power = 1:  G = sum pi_i x_i x_i^T)
//
// Why pairs.  The single-CTA kernel is bound by shared-memory bandwidth and by the tile loads, not by the tensor pipe
// (profiles/r02_tf32_stats.txt): a kind::tf32 instruction covers only K = 8 rows, so M128 x N256 reads 12 KiB of
// operands per 128 clk (96 of the SM's 128 B/clk) before the TMA fill and the in-place transform get any, and every
// CTA loads three 128-feature blocks per tile.  With cta_group::2 the two SMs of a TPC run ONE M = 256, N = 256
// instruction: each CTA supplies its 128 rows of A and only HALF of B (128 columns) from its own shared memory, the
// hardware shares the halves.  Per CTA and K step that is 8 KiB instead of 12, and a CTA loads / transforms one block
// per tile (diagonal pair: A half = B half) or two (off-diagonal) instead of three.
//
// Decomposition.  The features are cut into 256-wide pair blocks P (1 or 2 of them).  A pair of TYPE (P, Q), P <= Q,
// accumulates the 256 x 256 block G[P, Q]: CTA r holds rows 256 P + 128 r .. + 127 (TMEM lanes) x 256 columns.  Types:
// (0,0) for d <= 256; (0,0), (0,1), (1,1) for d <= 512.  Every type sweeps ALL row tiles, split over its own pairs
// (tile t of slot s: s, s + nslots, ...); the off-diagonal type has twice the load / transform work per tile and gets
// proportionally more pairs.  CTA r of a diagonal pair also owns the column sums of feature block 2 P + r.
//
// Per CTA the roles are those of gram_tf32.cu (TMA producer, three transform teams, two of them draining TMEM), with
// the cross-CTA edges of the protocol:
//   * transform warps of BOTH CTAs arrive (one elected lane per warp, release.cluster) on the LEADER's `ready` barrier;
//   * only the leader (cluster rank 0) issues tcgen05.mma.cta_group::2; tcgen05.commit ... multicast::cluster releases
//     the stage (`empty`) and publishes the chunk (`tfull`) in both CTAs;
//   * the accumulator warps of both CTAs arrive on the leader's `tempty` after reading their own TMEM back;
//   * cluster barriers after set-up and before TMEM is freed, and each producer drains its `empty` barriers before
//     leaving, so no multicast arrive can land in a CTA that has exited.
#include "tf32.cuh"

namespace {

using namespace tf32;

constexpr int kPairStages = 12;             // upper bound; the ring uses min(kPairStages, budget / stage bytes)
constexpr int kPairTypes = 3;
constexpr int kCoefBytes = 256;             // per stage: pi of the tile's 16 rows (128 B) + y (128 B)

struct PairParams {
  Tf32Params b;                 // w, y, n, d, power, chunks_per_flush, box3d, partials, err, wmax, stats
  int nbp;                      // 256-feature pair blocks: 1 or 2
  int ntypes;                   // 1 or 3
  int first[kPairTypes + 1];    // first pair of each type; first[ntypes] = number of pairs (grid = 2 x that)
};

__host__ __device__ __forceinline__ void pair_type(int t, int nbp, int& pa, int& pb) {
  if (nbp == 1 || t == 0) {
    pa = 0;
    pb = 0;
  } else if (t == 1) {
    pa = 0;
    pb = 1;
  } else {
    pa = 1;
    pb = 1;
  }
}

// ---- cluster-scope PTX ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): the data the waiter needs are
  // this CTA's shared-memory writes, published to the tensor core by fence.proxy.async; a release.cluster arrive costs
  // ~1100 clk per call (measured, profiles/r02_tf32_pair_ablation.txt)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool wait_or_abort_cluster(uint64_t* bar, uint32_t parity, unsigned int* err) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
  const long long t0 = clock64();
  unsigned int spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 255u) == 0u) {
      if (clock64() - t0 > 6000000000LL) {
        atomicExch(err, 1u);
        return false;
      }
      if (*reinterpret_cast<volatile unsigned int*>(err) != 0u) return false;
    }
  }
  return true;
}
__device__ __forceinline__ void tc2_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}
__device__ __forceinline__ void tc2_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor of the pair instruction: as umma_idesc, M = 256
__device__ __forceinline__ uint32_t umma_idesc_pair(int n_cols) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (uint32_t(n_cols >> 3) << 17) |
         (uint32_t(256 >> 4) << 24);
}

struct PairCtx {
  unsigned char* smem;
  uint64_t *full_bar, *empty_bar, *tfull_bar;
  uint32_t ready_remote, tempty_remote;    // the LEADER's `ready[0]` / `tempty[0]` as shared::cluster addresses
  float* red_all;
  double* dred_all;
  uint32_t tmem_base;
  int stage_bytes, nst, nfb, slot, nslots, team;
  int my_tiles, my_chunks;
  bool own_s0;
  bool coef_bulk;            // the producer copies the tile's pi (and y) into the stage's coefficient slot
  uint32_t coef_off;         // byte offset of the coefficient slots (kCoefBytes per stage) behind the stages
};

// One transform team (four warps): tiles team, team + 3, ... of the CTA.  IS_ACC teams (1, 2) own the level-2
// accumulators of column half team - 1.  HAS_SUMS (diagonal pairs): the column sums of the CTA's feature block ride along.
template <int NSPLIT, bool HAS_Y, bool IS_ACC, bool HAS_SUMS>
__device__ __forceinline__ void pair_team_body(const Tf32Params& p, const PairCtx& cx) {
  unsigned char* smem = cx.smem;
  uint64_t *full_bar = cx.full_bar, *tfull_bar = cx.tfull_bar, *empty_bar = cx.empty_bar;
  const uint32_t tmem_base = cx.tmem_base;
  const int stage_bytes = cx.stage_bytes, nst = cx.nst, nfb = cx.nfb, slot = cx.slot, team = cx.team;
  const int my_tiles = cx.my_tiles, my_chunks = cx.my_chunks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tt = threadIdx.x & 127;                      // thread of the team
  const int q = tt & 7, rr = tt >> 3;                    // logical 16-byte unit, row of the tile
  // SWIZZLE_128B_ATOM_32B: the 32-byte unit q >> 1 of row rr sits at unit (q >> 1) ^ (rr & 3)
  const uint32_t off = uint32_t(rr * 128 + (((((q >> 1) ^ (rr & 3)) << 1) | (q & 1)) << 4));
  const uint32_t lo_off = uint32_t(nfb * kBlkBytes);
  const bool own_s0 = cx.own_s0 && (q == 0);
  float* red = cx.red_all + team * 1024;
  double* dred = cx.dred_all + team * 32;
  const int bar_id = 1 + team;
  float s1acc[HAS_SUMS ? 4 : 1][4], syacc[(HAS_SUMS && HAS_Y) ? 4 : 1][4];
#pragma unroll
  for (int c = 0; c < (HAS_SUMS ? 4 : 1); ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) s1acc[c][k] = 0.f;
#pragma unroll
  for (int c = 0; c < ((HAS_SUMS && HAS_Y) ? 4 : 1); ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) syacc[c][k] = 0.f;
  double s1d = 0.0, syd = 0.0, s0 = 0.0, swy = 0.0;
  const double wscale = ldexp(1.0, -weight_exponent(p.wmax));      // exact power of two

  // ---- accumulator state (teams 1, 2: column half h = team - 1; thread = TMEM lane quarter * 32 + lane) ----
  const int h = team - 1;
  const int quarter = warp & 3;                          // hardware: a warp reaches TMEM lanes 32 (warp id % 4) ..
  float acc[IS_ACC ? 128 : 2];
#pragma unroll
  for (int i = 0; i < (IS_ACC ? 128 : 2); ++i) acc[i] = 0.f;
  double* gp = p.gpart64 + ((size_t(blockIdx.x) * 2 + (IS_ACC ? h : 0)) * 128 + size_t(quarter * 32 + lane)) * 128;
  bool flushed = false;
  int next_drain = 0;
  bool ok = true;

  auto flush_acc = [&]() {
    if (IS_ACC) {
#pragma unroll
      for (int i = 0; i < (IS_ACC ? 128 : 2); i += 2) {
        double2 v = make_double2(double(acc[i]), double(acc[i + 1]));
        if (flushed) {
          const double2 o = *reinterpret_cast<const double2*>(gp + i);
          v.x += o.x;
          v.y += o.y;
        }
        *reinterpret_cast<double2*>(gp + i) = v;
        acc[i] = 0.f;
        acc[i + 1] = 0.f;
      }
    }
    flushed = true;
  };
  auto drain = [&](int ch) -> bool {
    const int buf = ch & 1;
    if (!wait_or_abort(&tfull_bar[buf], uint32_t(ch >> 1) & 1u, p.err)) return false;
    tc_fence_after();
    if (IS_ACC && !(p.window & 4)) {
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * 256 + h * 128);
#pragma unroll
      for (int c = 0; c < (IS_ACC ? 8 : 0); ++c) {
        float v[16];
        tc_ld16(taddr + uint32_t(c * 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[(c * 16 + j) % (IS_ACC ? 128 : 2)] += v[j];
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(cx.tempty_remote + uint32_t(buf) * 8u);
    if (((ch + 1) % p.chunks_per_flush) == 0) flush_acc();
    return true;
  };

  auto flush_sums = [&]() {
    if (HAS_SUMS) {
      const int wq = tt >> 5;
#pragma unroll
      for (int c = 0; c < (HAS_SUMS ? 4 : 1); ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v = s1acc[c][k];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (lane < 8) red[(wq * 2 + 0) * 128 + c * 32 + lane * 4 + k] = v;
          s1acc[c][k] = 0.f;
          if (HAS_Y) {
            float u = syacc[(HAS_SUMS && HAS_Y) ? c : 0][k];
            u += __shfl_xor_sync(0xffffffffu, u, 8);
            u += __shfl_xor_sync(0xffffffffu, u, 16);
            if (lane < 8) red[(wq * 2 + 1) * 128 + c * 32 + lane * 4 + k] = u;
            syacc[(HAS_SUMS && HAS_Y) ? c : 0][k] = 0.f;
          }
        }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      s1d += double((red[0 * 128 + tt] + red[2 * 128 + tt]) + (red[4 * 128 + tt] + red[6 * 128 + tt]));
      if (HAS_Y) syd += double((red[1 * 128 + tt] + red[3 * 128 + tt]) + (red[5 * 128 + tt] + red[7 * 128 + tt]));
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
  };

  // Row coefficients: the producer's bulk copy delivers pi (and y) of a full tile with the tile itself (no global-load
  // latency in this loop: it was THE bound of the single-CTA kernel, profiles/r02_tf32_pair_ablation.txt); only a
  // ragged last tile, or unaligned pi / y, is read from global memory here.
  auto load_row = [&](int it, double& pid, double& yd) {
    const int64_t row = (int64_t(slot) + int64_t(it) * cx.nslots) * kR + rr;
    pid = (it < my_tiles && row < p.n) ? p.w[row] : 0.0;
    yd = (HAS_Y && it < my_tiles && row < p.n) ? p.y[row] : 0.0;
  };

  int done = 0;
  long long t_full = 0, t_busy = 0, t_drain = 0, t_a = 0, t_b = 0, t_c = 0, t_d = 0, t_e = 0;
  int s = team % nst;
  uint32_t ph = uint32_t(team / nst) & 1u;
  for (int it = team; it < my_tiles; it += kTeams) {
    const bool in_smem = cx.coef_bulk && ((int64_t(slot) + int64_t(it) * cx.nslots + 1) * kR <= p.n);
    double pid = 0.0, yd = 0.0;
    if (!in_smem) load_row(it, pid, yd);
    const long long k0 = p.stats ? clock64() : 0;
    if (IS_ACC) {                                        // chunk c - 2 is complete by now: drain it before chunk c
      const int ch = it / kTpc;
      while (ok && next_drain + 2 <= ch) ok = drain(next_drain++);
      if (!ok) break;
    }
    const long long k1 = p.stats ? clock64() : 0;
    // see gram_tf32.cu: `empty` one phase back first, so that `full` cannot be mistaken for the previous phase
    ok = wait_or_abort(&empty_bar[s], ph ^ 1u, p.err) && wait_or_abort(&full_bar[s], ph, p.err);
    if (!ok) break;
    const long long k2 = p.stats ? clock64() : 0;
    t_drain += k1 - k0;
    t_full += k2 - k1;
    unsigned char* sb = smem + size_t(s) * stage_bytes;
    if (in_smem) {
      const double* cw = reinterpret_cast<const double*>(smem + cx.coef_off + uint32_t(s) * kCoefBytes);
      pid = cw[rr];
      if (HAS_Y) yd = cw[kR + rr];
    }
    const double wd = (p.power == 2) ? pid * pid : pid;
    const double pis = pid * wscale;                       // normalised weight, <= 1
    const float sc = (p.power == 2) ? float(pis) : float(sqrt(pis));
    const float c1 = float(pis);
    const float cy = HAS_Y ? float(((p.power == 2) ? pis * pis : pis) * yd) : 0.f;
    if (own_s0) {
      s0 += wd;
      if (HAS_Y) swy = fma(wd, yd, swy);
    }
    const long long k3 = p.stats ? clock64() : 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i < nfb && !(p.window & 2)) {
        const bool sums = HAS_SUMS && (i == 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4* ptr = reinterpret_cast<float4*>(sb + i * kBlkBytes + c * kChunkBytes + off);
          const float4 x = *ptr;
          if (sums) {
            s1acc[HAS_SUMS ? c : 0][0] = fmaf(c1, x.x, s1acc[HAS_SUMS ? c : 0][0]);
            s1acc[HAS_SUMS ? c : 0][1] = fmaf(c1, x.y, s1acc[HAS_SUMS ? c : 0][1]);
            s1acc[HAS_SUMS ? c : 0][2] = fmaf(c1, x.z, s1acc[HAS_SUMS ? c : 0][2]);
            s1acc[HAS_SUMS ? c : 0][3] = fmaf(c1, x.w, s1acc[HAS_SUMS ? c : 0][3]);
            if (HAS_Y) {
              syacc[(HAS_SUMS && HAS_Y) ? c : 0][0] = fmaf(cy, x.x, syacc[(HAS_SUMS && HAS_Y) ? c : 0][0]);
              syacc[(HAS_SUMS && HAS_Y) ? c : 0][1] = fmaf(cy, x.y, syacc[(HAS_SUMS && HAS_Y) ? c : 0][1]);
              syacc[(HAS_SUMS && HAS_Y) ? c : 0][2] = fmaf(cy, x.z, syacc[(HAS_SUMS && HAS_Y) ? c : 0][2]);
              syacc[(HAS_SUMS && HAS_Y) ? c : 0][3] = fmaf(cy, x.w, syacc[(HAS_SUMS && HAS_Y) ? c : 0][3]);
            }
          }
          const float z0 = sc * x.x, z1 = sc * x.y, z2 = sc * x.z, z3 = sc * x.w;
          const uint32_t h0 = to_tf32(z0), h1 = to_tf32(z1), h2 = to_tf32(z2), h3 = to_tf32(z3);
          *reinterpret_cast<uint4*>(ptr) = make_uint4(h0, h1, h2, h3);
          if (NSPLIT == 3) {
            const uint32_t l0 = to_tf32(z0 - __uint_as_float(h0)), l1 = to_tf32(z1 - __uint_as_float(h1)),
                           l2 = to_tf32(z2 - __uint_as_float(h2)), l3 = to_tf32(z3 - __uint_as_float(h3));
            *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(ptr) + lo_off) = make_uint4(l0, l1, l2, l3);
          }
        }
      }
    }
    const long long k4 = p.stats ? clock64() : 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core
    const long long k5 = p.stats ? clock64() : 0;
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(cx.ready_remote + uint32_t(s) * 8u);   // the leader's barrier: 4 + 4 warps per tile
    const long long k6 = p.stats ? clock64() : 0;
    if ((++done % kFlushTiles) == 0) flush_sums();
    if (p.stats) {
      const long long k7 = clock64();
      t_a += k3 - k2;
      t_b += k4 - k3;
      t_c += k5 - k4;
      t_d += k6 - k5;
      t_e += k7 - k6;
    }
    s += kTeams;                                           // nst >= 4 > kTeams: at most one wrap
    if (s >= nst) {
      s -= nst;
      ph ^= 1u;
    }
    if (p.stats) t_busy += clock64() - k2;
  }
  if (p.stats && tt == 0 && team < 2) {
    p.stats[size_t(blockIdx.x) * 8 + (team == 0 ? 1 : 6)] = (team == 0) ? t_full : t_drain;
    if (team == 0) p.stats[size_t(blockIdx.x) * 8 + 2] = t_busy;
    if (team == 0) {
      long long* x = p.stats + size_t(gridDim.x) * 8 + size_t(blockIdx.x) * 8;
      x[0] = t_a;
      x[1] = t_b;
      x[2] = t_c;
      x[3] = t_d;
      x[4] = t_e;
    }
  }
  if (ok) {
    flush_sums();
    if (HAS_SUMS) {
      p.spart[((size_t(blockIdx.x) * kTeams + team) * 2 + 0) * 128 + tt] = s1d;
      p.spart[((size_t(blockIdx.x) * kTeams + team) * 2 + 1) * 128 + tt] = syd;
    }
    if (cx.own_s0) {
      if (q == 0) {
        dred[rr * 2 + 0] = s0;
        dred[rr * 2 + 1] = swy;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (tt == 0) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < kR; ++r) {
          a += dred[r * 2 + 0];
          b += dred[r * 2 + 1];
        }
        p.s0part[(size_t(blockIdx.x) * kTeams + team) * 2 + 0] = a;
        p.s0part[(size_t(blockIdx.x) * kTeams + team) * 2 + 1] = b;
      }
    }
    if (IS_ACC) {
      while (ok && next_drain < my_chunks) ok = drain(next_drain++);
      if (ok) flush_acc();
    }
  }
}

template <int NSPLIT, bool HAS_Y>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    gram_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmap2, const __grid_constant__ CUtensorMap tmap3,
                          const PairParams pp) {
  const Tf32Params& p = pp.b;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);   // same offset in both CTAs of the pair
  const uint32_t smem_base = smem_u32(smem);
  unsigned char* tail = smem + kSmemBudget;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);          // [kPairStages] TMA -> transform (own CTA)
  uint64_t* ready_bar = full_bar + kPairStages;                     // [kPairStages] transform (both CTAs) -> MMA (leader)
  uint64_t* empty_bar = ready_bar + kPairStages;                    // [kPairStages] MMA -> TMA (multicast to both)
  uint64_t* tfull_bar = empty_bar + kPairStages;                    // [2] MMA -> accumulator warps (multicast)
  uint64_t* tempty_bar = tfull_bar + 2;                             // [2] accumulator warps (both CTAs) -> MMA (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* red_all = reinterpret_cast<float*>(tail + 512);            // [kTeams][4 warps][2][128] column-sum exchange
  double* dred_all = reinterpret_cast<double*>(tail + 512 + kTeams * 4096);   // [kTeams][16][2] S0 / Swy exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  int type = 0;
  while (type + 1 < pp.ntypes && pair >= pp.first[type + 1]) ++type;
  const int slot = pair - pp.first[type];
  const int nslots = pp.first[type + 1] - pp.first[type];
  int pa, pb;
  pair_type(type, pp.nbp, pa, pb);
  const bool diag = (pa == pb);
  const int fb_a = 2 * pa + int(rank), fb_b = 2 * pb + int(rank);
  const int nfb = diag ? 1 : 2;
  const uint32_t a_off = 0u, b_off = diag ? 0u : uint32_t(kBlkBytes);
  const int stage_bytes = nfb * kBlkBytes * (NSPLIT == 3 ? 2 : 1);
  int nst = (kSmemBudget - kPairStages * kCoefBytes) / stage_bytes;
  if (nst > kPairStages) nst = kPairStages;
  const uint32_t coef_off = uint32_t(nst) * uint32_t(stage_bytes);      // coefficient slots behind the stages
  const bool coef_bulk = ((reinterpret_cast<uintptr_t>(p.w) | reinterpret_cast<uintptr_t>(p.y)) & 15u) == 0u;
  const int ntiles = int((p.n + kR - 1) / kR);
  const int my_tiles = (ntiles > slot) ? (ntiles - slot + nslots - 1) / nslots : 0;
  const int my_chunks = (my_tiles + kTpc - 1) / kTpc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 8);      // 4 transform warps of each CTA per tile
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 16);    // 8 accumulator warps of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM of both SMs: all 512 columns (two 256-column accumulator buffers)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp < 4) {
    // ===== warpgroup 0: TMA producer (warp 0, both CTAs) and MMA issuer (warp 1 of the leader) ====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long t_wait = 0;
      const long long t_begin = clock64();
      bool ok = true;
      for (int it = 0; it < my_tiles; ++it) {
        const long long c0 = clock64();
        if (!wait_or_abort(&empty_bar[s], ph ^ 1u, p.err)) {
          ok = false;
          break;
        }
        t_wait += clock64() - c0;
        const int row0 = (slot + it * nslots) * kR;
        const uint32_t sb = smem_base + uint32_t(s) * uint32_t(stage_bytes);
        const bool coef = coef_bulk && (int64_t(row0) + kR <= p.n);
        const uint32_t coef_tx = coef ? uint32_t(kR * 8 * (HAS_Y ? 2 : 1)) : 0u;
        mbar_arrive_expect_tx(&full_bar[s], ((p.window & 8) ? 0u : uint32_t(nfb * kBlkBytes)) + coef_tx);
        if (coef) {
          unsigned char* cdst = smem + coef_off + uint32_t(s) * kCoefBytes;
          bulk_g2s(cdst, p.w + row0, kR * 8, &full_bar[s]);
          if (HAS_Y) bulk_g2s(cdst + kR * 8, p.y + row0, kR * 8, &full_bar[s]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i < nfb && !(p.window & 8)) {
            const int fbi = (i == 0) ? fb_a : fb_b;
            if (p.box3d) {       // one box: 32 floats x 16 rows x 4 column groups
              tma_load_3d_f32(sb + uint32_t(i * kBlkBytes), &tmap3, 0, row0, fbi * 4, &full_bar[s]);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                tma_load_2d_f32(sb + uint32_t(i * kBlkBytes + c * kChunkBytes), &tmap2, fbi * kMB + c * 32, row0,
                                &full_bar[s]);
            }
          }
        }
        if (++s == nst) {
          s = 0;
          ph ^= 1u;
        }
      }
      // tail: every release of a stage this CTA filled must have landed before the CTA may exit (multicast arrives
      // from the leader's tcgen05.commit target this CTA's shared memory)
      for (int k = 0; ok && k < nst && k < my_tiles; ++k) {
        if (!wait_or_abort(&empty_bar[s], ph ^ 1u, p.err)) break;
        if (++s == nst) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (p.stats) {
        p.stats[size_t(blockIdx.x) * 8 + 0] = t_wait;
        p.stats[size_t(blockIdx.x) * 8 + 7] = clock64() - t_begin;
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_pair(256);
      const uint32_t lo_off = uint32_t(nfb * kBlkBytes);   // Z_lo blocks follow the Z_hi blocks of a stage
      const uint64_t dconst = umma_desc(0);
      auto desc = [&](uint32_t addr) { return dconst | uint64_t((addr & 0x3FFFFu) >> 4); };
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      long long t_ready = 0, t_tempty = 0, t_issue = 0;
      for (int it = 0; it < my_tiles && ok; ++it) {
        const int ch = it / kTpc;
        const int tin = it - ch * kTpc;
        const int buf = ch & 1;
        const long long c0 = clock64();
        if (tin == 0) {
          ok = wait_or_abort_cluster(&tempty_bar[buf], (uint32_t(ch >> 1) & 1u) ^ 1u, p.err);
          if (!ok) break;
          tc_fence_after();
        }
        const long long c1 = clock64();
        ok = wait_or_abort_cluster(&ready_bar[s], ph, p.err);
        if (!ok) break;
        tc_fence_after();
        const long long c2 = clock64();
        t_tempty += c1 - c0;
        t_ready += c2 - c1;
        const uint32_t sb = smem_base + uint32_t(s) * uint32_t(stage_bytes);
        const uint32_t dcol = tmem_base + uint32_t(buf * 256);
#pragma unroll
        for (int ks = 0; ks < kR / 8; ++ks) {
          if (p.window & 1) break;
          const uint32_t first = (tin == 0 && ks == 0) ? 0u : 1u;
          const uint32_t kb = sb + uint32_t(ks * 1024);
          const uint64_t a_hi = desc(kb + a_off), b_hi = desc(kb + b_off);
          tc2_mma_tf32(dcol, a_hi, b_hi, idesc, first);
          if (NSPLIT == 3) {
            tc2_mma_tf32(dcol, desc(kb + lo_off + a_off), b_hi, idesc, 1u);
            tc2_mma_tf32(dcol, a_hi, desc(kb + lo_off + b_off), idesc, 1u);
          }
        }
        tc2_commit_both(&empty_bar[s]);                                        // stage free in both CTAs
        if (tin == kTpc - 1 || it == my_tiles - 1) tc2_commit_both(&tfull_bar[buf]);   // chunk complete in both TMEMs
        t_issue += clock64() - c2;
        if (++s == nst) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (p.stats) {
        p.stats[size_t(blockIdx.x) * 8 + 3] = t_ready;
        p.stats[size_t(blockIdx.x) * 8 + 4] = t_tempty;
        p.stats[size_t(blockIdx.x) * 8 + 5] = t_issue;
      }
    }
  } else {
    // ===== warpgroups 1-3: the transform teams; teams 1 and 2 also hold the level-2 accumulators ================
    PairCtx cx;
    cx.smem = smem;
    cx.full_bar = full_bar;
    cx.empty_bar = empty_bar;
    cx.tfull_bar = tfull_bar;
    cx.ready_remote = map_to_cta(smem_u32(ready_bar), 0u);
    cx.tempty_remote = map_to_cta(smem_u32(tempty_bar), 0u);
    cx.red_all = red_all;
    cx.dred_all = dred_all;
    cx.tmem_base = tmem_base;
    cx.stage_bytes = stage_bytes;
    cx.nst = nst;
    cx.nfb = nfb;
    cx.slot = slot;
    cx.nslots = nslots;
    cx.team = (warp >> 2) - 1;
    cx.my_tiles = my_tiles;
    cx.my_chunks = my_chunks;
    cx.own_s0 = (type == 0 && rank == 0);
    cx.coef_bulk = coef_bulk;
    cx.coef_off = coef_off;
    if (cx.team == 0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
      if (diag) pair_team_body<NSPLIT, HAS_Y, false, true>(p, cx);
      else pair_team_body<NSPLIT, HAS_Y, false, false>(p, cx);
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
      if (diag) pair_team_body<NSPLIT, HAS_Y, true, true>(p, cx);
      else pair_team_body<NSPLIT, HAS_Y, true, false>(p, cx);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // both CTAs are done with each other's shared memory, barriers and TMEM
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Sum the per-CTA partials in slot order, mirror the upper triangle, write [S0, Swy, S1, Sy, G] (FP64).
__global__ void __launch_bounds__(256) gram_tf32_pair_finalize_kernel(const PairParams pp, int has_y, int want_gram,
                                                                      double* out) {
  const Tf32Params& p = pp.b;
  const int d = p.d, nb = 2 * pp.nbp;
  const int npairs_blk = nb * (nb + 1) / 2;
  const int64_t gtotal = int64_t(npairs_blk) * kMB * kMB;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool bad = *p.err != 0u;
  const int wexp = weight_exponent(p.wmax);
  const double up1 = ldexp(1.0, wexp), upw = ldexp(1.0, wexp * p.power);   // undo the weight normalisation
  if (idx < gtotal) {
    if (!want_gram) return;
    int pidx = int(idx / (kMB * kMB));
    const int i = int((idx / kMB) % kMB), j = int(idx % kMB);     // G block (a, b), a <= b, element (i, j)
    int a = 0;
    while (pidx >= nb - a) {
      pidx -= nb - a;
      ++a;
    }
    const int b = a + pidx;
    const int fi = a * kMB + i, fj = b * kMB + j;
    if (fi >= d || fj >= d) return;
    if (a == b && i > j) return;
    const int pa = a >> 1, r = a & 1, pb = b >> 1, h = b & 1;
    const int type = (pp.nbp == 1) ? 0 : (pa == 0 ? pb : 2);
    double s = 0.0;
    for (int sl = pp.first[type]; sl < pp.first[type + 1]; ++sl)
      s += p.gpart64[((size_t(sl) * 2 + r) * 2 + h) * (kMB * kMB) + size_t(i) * kMB + j];
    s *= upw;
    if (bad) s = nan("");
    double* G = out + 2 + 2 * d;
    G[size_t(fi) * d + fj] = s;
    G[size_t(fj) * d + fi] = s;
    return;
  }
  const int64_t k = idx - gtotal;
  if (k < 2) {          // S0, Swy: CTA 0 of the pairs of type 0
    double s = 0.0;
    for (int sl = pp.first[0]; sl < pp.first[1]; ++sl)
      for (int t = 0; t < kTeams; ++t) s += p.s0part[((size_t(sl) * 2) * kTeams + t) * 2 + k];
    out[k] = bad ? nan("") : ((k == 1 && !has_y) ? 0.0 : s);
    return;
  }
  const int64_t f2 = k - 2;
  if (f2 < 2 * int64_t(d)) {
    const int which = int(f2 / d), f = int(f2 % d);
    const int blk = f / kMB, fin = f % kMB;
    const int type = (pp.nbp == 1) ? 0 : ((blk >> 1) == 0 ? 0 : 2);     // the diagonal pair of the block
    const int r = blk & 1;
    double s = 0.0;
    if (which == 0 || has_y)
      for (int sl = pp.first[type]; sl < pp.first[type + 1]; ++sl)
        for (int t = 0; t < kTeams; ++t) s += p.spart[(((size_t(sl) * 2 + r) * kTeams + t) * 2 + which) * 128 + fin];
    out[2 + which * d + f] = bad ? nan("") : s * (which == 0 ? up1 : upw);
  }
}

}  // namespace

// Host side of the pair path.  `p` arrives with w, y, n, d, power, box3d, err, wmax filled in by
// rlvi_weighted_moments_f32 (which also ran the weight-maximum kernel); returns RLVI_ERR_UNSUPPORTED when the shape or
// the device does not take pairs (the caller then runs the single-CTA kernel).
int rlvi_tf32_pair_moments(rlvi_ctx* ctx, const CUtensorMap& tmap, const CUtensorMap& tmap3, tf32::Tf32Params p,
                           int precision, int want_gram, double* out, cudaStream_t st) {
  const int nb = (p.d + kMB - 1) / kMB;
  if (nb != 2 && nb != 4) return RLVI_ERR_UNSUPPORTED;
  if (getenv("RLVI_TF32_NO_PAIR")) return RLVI_ERR_UNSUPPORTED;
  PairParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.nbp = nb / 2;
  pp.ntypes = (pp.nbp == 1) ? 1 : 3;

  const void* fn;
  if (precision == RLVI_TF32X1) fn = p.y ? (const void*)gram_tf32_pair_kernel<1, true> : (const void*)gram_tf32_pair_kernel<1, false>;
  else fn = p.y ? (const void*)gram_tf32_pair_kernel<3, true> : (const void*)gram_tf32_pair_kernel<3, false>;
  RLVI_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));

  // how many pairs can be resident at once (one CTA per SM, two SMs of a TPC per pair)
  int max_pairs = 0;
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(unsigned(ctx->sm_count / 2 * 2), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemTotal;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    // pairs are independent of each other: a wrong answer here costs speed (a second wave), never correctness
    if (cudaOccupancyMaxActiveClusters(&max_pairs, fn, &cfg) != cudaSuccess || max_pairs < 1) {
      cudaGetLastError();
      max_pairs = ctx->sm_count / 2;
    }
  }
  if (max_pairs > ctx->sm_count / 2) max_pairs = ctx->sm_count / 2;
  if (getenv("RLVI_TF32_STATS")) fprintf(stderr, "[tf32 pair] resident pairs: %d\n", max_pairs);
  if (const char* e = getenv("RLVI_TF32_PAIRS")) {
    const int v = atoi(e);
    if (v > 0 && v < max_pairs) max_pairs = v;
  }
  const int64_t ntiles = (p.n + kR - 1) / kR;
  if (max_pairs < pp.ntypes) return RLVI_ERR_UNSUPPORTED;

  // pairs per type: in proportion to the per-tile cost (off-diagonal pairs load and transform two blocks per tile)
  int count[kPairTypes] = {0, 0, 0};
  if (pp.ntypes == 1) {
    count[0] = int(ntiles < max_pairs ? ntiles : max_pairs);
  } else {
    double wdiag = (precision == RLVI_TF32X1) ? 1.0 : 1.0, woff = (precision == RLVI_TF32X1) ? 1.6 : 1.15;
    if (const char* e = getenv("RLVI_TF32_PAIR_OFFDIAG")) woff = atof(e);
    int total = max_pairs;
    if (ntiles * 3 < total) total = int(ntiles) * 3;
    int off = int(double(total) * woff / (2.0 * wdiag + woff) + 0.5);
    if (off < 1) off = 1;
    int dg = (total - off) / 2;
    if (dg < 1) dg = 1;
    off = total - 2 * dg;
    if (off < 1) return RLVI_ERR_UNSUPPORTED;
    count[0] = dg;
    count[1] = off;
    count[2] = dg;
  }
  pp.first[0] = 0;
  for (int t = 0; t < pp.ntypes; ++t) pp.first[t + 1] = pp.first[t] + count[t];
  for (int t = pp.ntypes; t < kPairTypes; ++t) pp.first[t + 1] = pp.first[pp.ntypes];
  const int npairs = pp.first[pp.ntypes];
  const int grid = 2 * npairs;

  p.nb = nb;
  p.ngroups = 0;
  p.nslots = 0;
  p.chunks_per_flush = 256;
  p.window = 0;
  if (const char* e = getenv("RLVI_TF32_PAIR_DEBUG")) p.window = atoi(e);   // bring-up: 1 no MMA, 2 no transform, 4 no TMEM read, 8 no TMA
  const size_t gbytes = size_t(grid) * 2 * kMB * kMB * sizeof(double);
  const size_t sbytes = size_t(grid) * kTeams * 2 * 128 * sizeof(double);
  const size_t s0bytes = size_t(grid) * kTeams * 2 * sizeof(double);
  void* scratch = nullptr;
  const int rc = rlvi_scratch(ctx, 4096 + gbytes + sbytes + s0bytes, &scratch);
  if (rc != RLVI_OK) return rc;
  char* base = static_cast<char*>(scratch);
  if (reinterpret_cast<unsigned int*>(base + 2048) != p.err) {
    rlvi_set_error("rlvi_tf32_pair_moments: the scratch moved after the weight-maximum kernel was queued");
    return RLVI_ERR_CUDA;
  }
  p.gpart64 = reinterpret_cast<double*>(base + 4096);
  p.spart = reinterpret_cast<double*>(base + 4096 + gbytes);
  p.s0part = reinterpret_cast<double*>(base + 4096 + gbytes + sbytes);
  p.progress = nullptr;
  p.stats = nullptr;
  if (getenv("RLVI_TF32_STATS")) {
    RLVI_CUDA(cudaMalloc(&p.stats, size_t(grid) * 128));
    RLVI_CUDA(cudaMemset(p.stats, 0, size_t(grid) * 128));
  }
  pp.b = p;

  if (precision == RLVI_TF32X1) {
    if (p.y) gram_tf32_pair_kernel<1, true><<<grid, kThreads, kSmemTotal, st>>>(tmap, tmap3, pp);
    else gram_tf32_pair_kernel<1, false><<<grid, kThreads, kSmemTotal, st>>>(tmap, tmap3, pp);
  } else {
    if (p.y) gram_tf32_pair_kernel<3, true><<<grid, kThreads, kSmemTotal, st>>>(tmap, tmap3, pp);
    else gram_tf32_pair_kernel<3, false><<<grid, kThreads, kSmemTotal, st>>>(tmap, tmap3, pp);
  }
  RLVI_LAUNCH_CHECK(ctx);
  if (p.stats) {     // bring-up only: synchronise and print the per-role cycle counters (mean over the CTAs of a type)
    RLVI_CUDA(cudaStreamSynchronize(st));
    long long* hst = static_cast<long long*>(malloc(size_t(grid) * 128));
    cudaMemcpy(hst, p.stats, size_t(grid) * 128, cudaMemcpyDeviceToHost);
    const char* names[8] = {"producer wait empty", "team0 wait full", "team0 busy", "mma wait ready", "mma wait tempty",
                            "mma issue", "team1 drain", "producer total"};
    for (int t = 0; t < pp.ntypes; ++t) {
      const int cnt = pp.first[t + 1] - pp.first[t];
      const double tiles = double(ntiles) / cnt;
      for (int r = 0; r < 2; ++r) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int sl = pp.first[t]; sl < pp.first[t + 1]; ++sl)
          for (int k = 0; k < 8; ++k) acc[k] += double(hst[(size_t(sl) * 2 + r) * 8 + k]) / cnt;
        fprintf(stderr, "[tf32 pair stats] type %d cta %d (%d pairs, %.0f tiles/CTA), cycles per tile:", t, r, cnt, tiles);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s=%.0f", names[k], acc[k] / tiles);
        const char* fine[5] = {"coef+math", "transform", "proxy fence", "syncwarp+arrive", "flush sums"};
        fprintf(stderr, " | team0:");
        for (int k = 0; k < 5; ++k) {
          double a2 = 0;
          for (int sl = pp.first[t]; sl < pp.first[t + 1]; ++sl) a2 += double(hst[size_t(grid) * 8 + (size_t(sl) * 2 + r) * 8 + k]) / cnt;
          fprintf(stderr, " %s=%.0f", fine[k], a2 / tiles);
        }
        fprintf(stderr, "\n");
      }
    }
    free(hst);
    cudaFree(p.stats);
  }
  const int64_t total = int64_t(nb) * (nb + 1) / 2 * kMB * kMB + 2 + 2 * int64_t(p.d);
  gram_tf32_pair_finalize_kernel<<<int((total + 255) / 256), 256, 0, st>>>(pp, p.y ? 1 : 0, want_gram, out);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}
