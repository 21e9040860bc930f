// K2 (MLP + LayerNorm + residual) on CTA PAIRS (tcgen05 cta_group::2), seventh version.
//
//   h_out = h + LayerNorm(W2 relu(W1' z + b1') + b2)
//
// gfx_umma4.cu / gfx_umma6.cu keep both weight images (128 KB) resident in every CTA; what is left
// holds three 32 KB tiles, and the life cycle of those tiles (z -> residual -> output, two HBM
// round trips each) sets the pace, not the tensor pipe or HBM (the time per
// node is the same for an L2-resident working set).  A CTA pair splits the B operands -- each CTA
// holds the half of W1 / W2 that produces half of the N columns, 64 KB -- and runs M = 256 MMAs
// over both CTAs' 128-row tiles.  Per CTA that leaves room for FIVE stage buffers, each cycling
// z tile -> residual tile -> output tile as in gfx_umma6.cu, so the loads of tiles t+1 .. t+4
// are in flight while tile t is computed.  With B split, an M = 256, N = 128, K = 16 MMA reads
// 6 KB of operands per CTA instead of 8 KB for the same arithmetic, which takes GEMM 1 off the
// shared-memory bandwidth limit (88 -> 64 cycles per MMA in gfx_umma4.cu's timeline).
//
// TMEM (512 columns per CTA, as gfx_umma6.cu): D1 [0,256) fp32, A2 [256,384) fp16 hidden
// activation (GEMM 2's A operand), D2 [384,512) fp32.
//
// Warps per CTA: 0-7 epilogue A, 8-15 epilogue B, 16 weights + MMA issuer (rank 0 issues for the
// pair; rank 1's thread relays "my z tile has landed" to rank 0), 17 z loader, 18 residual
// loader, 19 output store (all TMA).  Barriers that collect arrivals from both CTAs (peer z
// landed, A2 full, D2 empty, weights ready) live in rank 0 and are reached with mapa + a remote
// arrive; completions of the MMAs are multicast to both CTAs by tcgen05.commit.
#include <cstdlib>

#include "gfx_common.cuh"
#include "gfx_pair.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v7k {

constexpr int HID = kMlpHidden, H = HID / 2;
constexpr int kTileM = 128;
constexpr int kKbBytes = kTileM * 128;        // one K block of a tile: [128 x 64] fp16
constexpr int kTileBytes = 2 * kKbBytes;      // a whole [128 x 128] fp16 tile
constexpr int kWPiece = 64 * 128;             // 64 weight rows x 64 columns (one CTA's share)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kA2Col = 256, kD2Col = 384;
constexpr int kEpiBWarp0 = 8, kMmaWarp = 16, kLoadWarp = 17, kResWarp = 18, kStoreWarp = 19, kWarps = 20;
constexpr int kStages = 5;

enum Bar {
  kBarWLocal = 0, kBarWReady = 1, kBarD1aFull = 2, kBarD1bFull = 3, kBarA2aFull = 4, kBarA2bFull = 5,
  kBarD2Full = 6, kBarD2Empty = 7,
  kBarA1Full = 8,                           // [stage]      z landed in this CTA
  kBarA1Peer = kBarA1Full + kStages,        // [stage]      (rank 0) z landed in rank 1
  kBarA1Empty = kBarA1Peer + kStages,       // [stage]      GEMM 1 has consumed z (multicast)
  kBarStageFree = kBarA1Empty + kStages,    // [stage]      the store has read the output
  kBarRFull = kBarStageFree + kStages,      // [stage][2]   residual half landed
  kBarOReady = kBarRFull + 2 * kStages,     // [stage][2]   output half written
  kNumBars = kBarOReady + 2 * kStages
};

struct Smem {
  static constexpr int off_w1 = 0;                                   // [kb 2][half 2] x 8 KB
  static constexpr int off_w2 = off_w1 + 4 * kWPiece;                // [kb 4] x 8 KB
  static constexpr int off_a1 = off_w2 + 4 * kWPiece;                // 5 stages x 32 KB
  static constexpr int off_xs = off_a1 + kStages * kTileBytes;       // float2[2][128] partial sums
  static constexpr int off_bar = off_xs + 2 * kTileM * 8;
  static constexpr int off_tmem = off_bar + kNumBars * 8;
  static constexpr int total = off_tmem + 8;
};
static_assert(Smem::total <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct alignas(64) Maps {
  CUtensorMap z, res, out;   // [n, 128] fp16, box 64 x 128, SWIZZLE_128B
};

struct Consts {              // kernel parameters = constant bank: free ALU operands
  float b1[HID], b2[kHidden], g[kHidden], b[kHidden];
};

struct Args {
  const __half *w1_img, *w2_img;
  int64_t n;
  long long *trace;          // developer timeline (tools/k2_trace.py); null in production
};

__device__ __forceinline__ void trace_ev(const Args &p, uint32_t it, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && it < 64) {
    p.trace[it * 16 + ev] = clock64();
    if (ev == 2) {                       // wall-clock stamp next to it: effective SM clock
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      p.trace[it * 16 + 14] = (long long)ns;
    }
  }
}

__device__ __forceinline__ uint32_t relu_pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// D1[:, HALF*128 + CH*64 .. +64) -> bias + ReLU -> fp16 -> A2; tells rank 0's MMA issuer
template <int HALF, int CH>
__device__ __forceinline__ void epi_a(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      int lane, uint32_t leader_bar) {
  constexpr int col0 = HALF * H + CH * 64;
  mbar_wait(bar + (HALF ? kBarD1bFull : kBarD1aFull), ph);
  tc_fence_after();
  float v[64];
  tmem_ld32(trow + col0, v);
  tmem_ld32(trow + col0 + 32, v + 32);
  tmem_ld_wait();
  uint32_t pk[32];
#pragma unroll
  for (int j = 0; j < 32; ++j)
    pk[j] = relu_pack2(v[2 * j] + c.b1[col0 + 2 * j], v[2 * j + 1] + c.b1[col0 + 2 * j + 1]);
  tmem_st16(trow + kA2Col + col0 / 2, pk);
  tmem_st16(trow + kA2Col + col0 / 2 + 16, pk + 16);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(leader_bar);
}

// D2[:, CH*64 .. +64) -> + b2 -> LayerNorm (stats shared with the other column half) -> * g + b
// + residual, in place in the stage buffer that now holds the tile's residual rows
template <int CH>
__device__ __forceinline__ void epi_b(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      uint32_t st, uint32_t sph, int lane, int quad, float2 *xs,
                                      uint8_t *rs, uint32_t d2_empty) {
  constexpr int col0 = CH * 64;
  const int r = quad * 32 + lane;
  mbar_wait(bar + kBarD2Full, ph);
  tc_fence_after();
  float u[64];
  tmem_ld32(trow + kD2Col + col0, u);
  tmem_ld32(trow + kD2Col + col0 + 32, u + 32);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(d2_empty);         // accumulator is in registers now
  float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    u[j] += c.b2[col0 + j];
    s1[j & 1] += u[j];
    s2[j & 1] = fmaf(u[j], u[j], s2[j & 1]);
  }
  xs[CH * kTileM + r] = make_float2(s1[0] + s1[1], s2[0] + s2[1]);
  named_bar_sync(1 + quad, 64);
  const float2 other = xs[(CH ^ 1) * kTileM + r];
  named_bar_sync(1 + quad, 64);                          // xs may be rewritten for the next tile
  const float mean = (s1[0] + s1[1] + other.x) * (1.f / kHidden);
  const float var = fmaxf((s2[0] + s2[1] + other.y) * (1.f / kHidden) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nm = -mean * rstd;
  mbar_wait(bar + kBarRFull + 2 * st + CH, sph);
  uint8_t *rrow = rs + CH * kKbBytes + r * 128;
  const int rx = r & 7;
#pragma unroll
  for (int gi = 0; gi < 8; ++gi) {
    uint4 *cell = reinterpret_cast<uint4 *>(rrow + ((gi ^ rx) << 4));
    const uint4 raw = *cell;
    const __half2 *hp = reinterpret_cast<const __half2 *>(&raw);
    float o[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 rr = __half22float2(hp[q]);
      const int j = gi * 8 + 2 * q;
      o[2 * q] = fmaf(fmaf(u[j], rstd, nm), c.g[col0 + j], rr.x + c.b[col0 + j]);
      o[2 * q + 1] = fmaf(fmaf(u[j + 1], rstd, nm), c.g[col0 + j + 1], rr.y + c.b[col0 + j + 1]);
    }
    *cell = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar + kBarOReady + 2 * st + CH);
}

// WIDE: GEMM 1 as 8 MMAs of N = 256 (each CTA supplies W1' rows [128 rank, +128)) instead of
// two column halves of 8 MMAs with N = 128 (rows [128 half + 64 rank, +64)).
template <bool WIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWarps * 32, 1)
umma7_mlp_pair_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Consts c, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2, *a1s = smem + L::off_a1;
  float2 *xs = reinterpret_cast<float2 *>(smem + L::off_xs);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::off_bar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (p.trace != nullptr && tid == 0) {          // every CTA: wall-clock start (end below)
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.trace[1024 + 2 * blockIdx.x] = (long long)ns;
  }
  if (warp == kMmaWarp) {
    tmem_alloc2(tmem_slot, kTmemCols);
  } else if (tid == 0) {
    mbar_init(bar + kBarWLocal, 1);
    mbar_init(bar + kBarWReady, 2);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar + kBarA1Full + s, 1);
      mbar_init(bar + kBarA1Peer + s, 1);
      mbar_init(bar + kBarA1Empty + s, 1);
      mbar_init(bar + kBarStageFree + s, 1);
      for (int j = 0; j < 2; ++j) {
        mbar_init(bar + kBarRFull + 2 * s + j, 1);
        mbar_init(bar + kBarOReady + 2 * s + j, 4);
      }
    }
    mbar_init(bar + kBarD1aFull, 1);
    mbar_init(bar + kBarD1bFull, 1);
    mbar_init(bar + kBarA2aFull, 16);          // 8 epilogue-A warps of each CTA
    mbar_init(bar + kBarA2bFull, 16);
    mbar_init(bar + kBarD2Full, 1);
    mbar_init(bar + kBarD2Empty, 16);          // 8 epilogue-B warps of each CTA
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                          // both CTAs' barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n = int(p.n);
  const int tiles = (n + kTileM - 1) / kTileM;
  const int pairs = (tiles + 1) / 2;
  const int cluster_id = blockIdx.x >> 1, clusters = gridDim.x >> 1;
  auto leader = [&](int b) { return map_to_cta(smem_u32(bar + b), 0); };

  if (warp < kEpiBWarp0) {
    // ================= epilogue A: D1 -> bias + ReLU -> fp16 -> A2 (TMEM) =========
    const int quad = warp & 3, ch = warp >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16);
    const uint32_t a2a = leader(kBarA2aFull), a2b = leader(kBarA2bFull);
    uint32_t it = 0;
    for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
      const uint32_t ph = it & 1;
      if (ch == 0) {
        epi_a<0, 0>(c, trow, bar, ph, lane, a2a);
        if (tid == 0) trace_ev(p, it, 6);
        epi_a<1, 0>(c, trow, bar, ph, lane, a2b);
        if (tid == 0) trace_ev(p, it, 7);
      } else {
        epi_a<0, 1>(c, trow, bar, ph, lane, a2a);
        epi_a<1, 1>(c, trow, bar, ph, lane, a2b);
      }
    }
  } else if (warp < kMmaWarp) {
    // ===== epilogue B: D2 -> bias + LayerNorm, + residual, in place in the stage buffer =====
    const int quad = warp & 3, ch = (warp - kEpiBWarp0) >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16);
    const uint32_t d2e = leader(kBarD2Empty);
    uint32_t it = 0;
    for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
      const uint32_t st = it % kStages, sph = (it / kStages) & 1;
      if (ch == 0 && quad == 0 && lane == 0) trace_ev(p, it, 11);
      if (ch == 0)
        epi_b<0>(c, trow, bar, it & 1, st, sph, lane, quad, xs, a1s + st * kTileBytes, d2e);
      else
        epi_b<1>(c, trow, bar, it & 1, st, sph, lane, quad, xs, a1s + st * kTileBytes, d2e);
      if (ch == 0 && quad == 0 && lane == 0) trace_ev(p, it, 9);
    }
  } else if (warp == kMmaWarp) {
    // ============ weights (both CTAs), MMA issue (rank 0), z-landed relay (rank 1) ============
    if (lane == 0) {
      mbar_arrive_expect_tx(bar + kBarWLocal, 8 * kWPiece);
      const uint8_t *w1g = reinterpret_cast<const uint8_t *>(p.w1_img);
      const uint8_t *w2g = reinterpret_cast<const uint8_t *>(p.w2_img);
      for (int kb = 0; kb < 2; ++kb)
        for (int half = 0; half < 2; ++half)      // rows [128 half + 64 rank, +64) of K block kb
          bulk_g2s(w1s + (kb * 2 + half) * kWPiece,
                   w1g + kb * (HID * 128) +
                       (WIDE ? H * int(rank) + 64 * half : H * half + 64 * int(rank)) * 128,
                   kWPiece, bar + kBarWLocal);
      for (int kb = 0; kb < 4; ++kb)              // rows [64 rank, +64) of K block kb
        bulk_g2s(w2s + kb * kWPiece, w2g + kb * (kHidden * 128) + 64 * int(rank) * 128, kWPiece,
                 bar + kBarWLocal);
      mbar_wait_parked(bar + kBarWLocal, 0);
      mbar_arrive_cluster(leader(kBarWReady));
      if (rank == 0) {
        mbar_wait(bar + kBarWReady, 0);
        constexpr uint32_t idesc = idesc_f16(2 * kTileM, H);     // M = 256 over the pair, N = 128
        const uint32_t w1a = smem_u32(w1s), w2a = smem_u32(w2s);
        uint32_t it = 0;
        for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
          const uint32_t s = it % kStages, ph2 = (it / kStages) & 1, ph = it & 1;
          const uint32_t za = smem_u32(a1s) + s * kTileBytes;
          mbar_wait(bar + kBarA1Full + s, ph2);
          mbar_wait(bar + kBarA1Peer + s, ph2);
          tc_fence_after();
          trace_ev(p, it, 2);
          if (WIDE) {
            constexpr uint32_t idesc_w = idesc_f16(2 * kTileM, HID);   // N = 256
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const int kb = kk >> 2, k = kk & 3;
              const uint64_t da = smem_desc_sw128(za + kb * kKbBytes + k * 32);
              const uint64_t db = smem_desc_sw128(w1a + kb * 2 * kWPiece + k * 32);
              mma2_f16_ss(tmem, da, db, idesc_w, kk != 0);
            }
            mma2_commit(bar + kBarD1aFull);
            mma2_commit(bar + kBarD1bFull);
          } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const int kb = kk >> 2, k = kk & 3;
                const uint64_t da = smem_desc_sw128(za + kb * kKbBytes + k * 32);
                const uint64_t db = smem_desc_sw128(w1a + (kb * 2 + half) * kWPiece + k * 32);
                mma2_f16_ss(tmem + half * H, da, db, idesc, kk != 0);
              }
              mma2_commit(bar + (half ? kBarD1bFull : kBarD1aFull));
            }
          }
          mma2_commit(bar + kBarA1Empty + s);          // z consumed in both CTAs
          trace_ev(p, it, 3);
          mbar_wait(bar + kBarA2aFull, ph);
          trace_ev(p, it, 4);
          mbar_wait(bar + kBarD2Empty, ph ^ 1);
          tc_fence_after();
          trace_ev(p, it, 5);
#pragma unroll
          for (int kk = 0; kk < HID / 16; ++kk) {
            if (kk == H / 16) {
              mbar_wait(bar + kBarA2bFull, ph);
              tc_fence_after();
            }
            const uint64_t db = smem_desc_sw128(w2a + (kk >> 2) * kWPiece + (kk & 3) * 32);
            mma2_f16_ts(tmem + kD2Col, tmem + kA2Col + kk * 8, db, idesc, kk != 0);
          }
          mma2_commit(bar + kBarD2Full);
          trace_ev(p, it, 8);
        }
      } else {
        uint32_t it = 0;
        for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
          const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
          mbar_wait_parked(bar + kBarA1Full + s, ph2);
          mbar_arrive_cluster(leader(kBarA1Peer + s));
        }
      }
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    // ============================ z tile loader (TMA) ============================
    if (lane == 0) {
      prefetch_tmap(&maps.z);
      uint32_t it = 0;
      for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        uint8_t *a1 = a1s + s * kTileBytes;
        mbar_wait_parked(bar + kBarStageFree + s, ph2 ^ 1);  // the previous tenant's output has left
        trace_ev(p, it, 0);
        mbar_arrive_expect_tx(bar + kBarA1Full + s, kTileBytes);
        const int row0 = (2 * pair + int(rank)) * kTileM;
        tma_load_2d(a1, &maps.z, 0, row0, bar + kBarA1Full + s);
        tma_load_2d(a1 + kKbBytes, &maps.z, 64, row0, bar + kBarA1Full + s);
      }
    }
    __syncwarp();
  } else if (warp == kResWarp) {
    // ============ residual rows into the stage, once GEMM 1 has consumed its z (TMA) ===========
    if (lane == 0) {
      prefetch_tmap(&maps.res);
      uint32_t it = 0;
      for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        uint8_t *a1 = a1s + s * kTileBytes;
        const int row0 = (2 * pair + int(rank)) * kTileM;
        mbar_wait_parked(bar + kBarA1Empty + s, ph2);
        trace_ev(p, it, 1);
        for (int j = 0; j < 2; ++j) {
          mbar_arrive_expect_tx(bar + kBarRFull + 2 * s + j, kKbBytes);
          tma_load_2d(a1 + j * kKbBytes, &maps.res, j * 64, row0, bar + kBarRFull + 2 * s + j);
        }
      }
    }
    __syncwarp();
  } else {
    // ============================ output store (TMA) ==============================
    if (lane == 0) {
      prefetch_tmap(&maps.out);
      uint32_t it = 0;
      for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        const uint8_t *a1 = a1s + s * kTileBytes;
        const int row0 = (2 * pair + int(rank)) * kTileM;
        for (int j = 0; j < 2; ++j) {
          mbar_wait_parked(bar + kBarOReady + 2 * s + j, ph2);
          if (row0 < n) tma_store_2d(&maps.out, j * 64, row0, a1 + j * kKbBytes);
        }
        bulk_commit();
        bulk_wait_read<0>();            // shared memory has been read: the stage takes its next z
        mbar_arrive(bar + kBarStageFree + s);
        trace_ev(p, it, 10);
      }
      bulk_wait_all();                  // every store has landed before the CTA exits
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                   // the peer may still be arriving on our barriers
  if (warp == kMmaWarp) tmem_dealloc2(tmem, kTmemCols);
  if (p.trace != nullptr && tid == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.trace[1024 + 2 * blockIdx.x + 1] = (long long)ns;
  }
}

}  // namespace v7k

extern long long *g_k2_trace;           // gfx_umma6.cu (gfx_debug_k2_trace)

int umma7_mlp_ln_residual(const gfx_model *m, int layer, const __half *z, const __half *h,
                          int64_t n, __half *h_out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) |
       reinterpret_cast<uintptr_t>(h_out)) & 15)
    return fail(GFX_ERR_ARGUMENT, "tcgen05 MLP: activation buffers must be 16-byte aligned");
  if (n > (int64_t(1) << 30))
    return fail(GFX_ERR_UNSUPPORTED, "tcgen05 MLP (CTA pairs): at most 2^30 nodes per call");
  v7k::Maps maps;
  int rc = tma::make_rows128_map(&maps.z, z, n, v7k::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.res, h, n, v7k::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.out, h_out, n, v7k::kTileM);
  if (rc) return rc;
  v7k::Consts c;
  const gfx_host_vectors &hv = m->host;
  for (int i = 0; i < kMlpHidden; ++i) c.b1[i] = hv.b1[size_t(layer) * kMlpHidden + i];
  for (int i = 0; i < kHidden; ++i) {
    c.b2[i] = hv.b2[size_t(layer) * kHidden + i];
    c.g[i] = hv.ln_g[size_t(layer) * kHidden + i];
    c.b[i] = hv.ln_b[size_t(layer) * kHidden + i];
  }
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  v7k::Args a{};
  a.w1_img = m->w1_img + wi; a.w2_img = m->w2_img + wi;
  a.n = n;
  a.trace = g_k2_trace;
  static const bool wide = [] {
    const char *v = getenv("GFX_K2_PAIR_WIDE");    // developer switch: GEMM 1 as N = 256 MMAs
    return !(v && *v == '0');
  }();
  auto kernel = wide ? v7k::umma7_mlp_pair_kernel<true> : v7k::umma7_mlp_pair_kernel<false>;
  GFX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v7k::Smem::total));
  const int64_t tiles = (n + v7k::kTileM - 1) / v7k::kTileM;
  const int64_t pairs = (tiles + 1) / 2;
  // As many CTA pairs as this GPU can hold at once (which SMs are fused off differs from chip to
  // chip, and a pair that does not fit would wait for a whole wave to finish).
  static int resident[64] = {};                   // per device; 0 = not asked yet
  int device = 0;
  GFX_CUDA(cudaGetDevice(&device));
  if (device >= 0 && device < 64 && resident[device] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs, 1, 1);
    cfg.blockDim = dim3(v7k::kWarps * 32, 1, 1);
    cfg.dynamicSmemBytes = v7k::Smem::total;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess ||
        max_clusters < 1) {
      (void)cudaGetLastError();
      max_clusters = kNumSMs / 2;
    }
    resident[device] = max_clusters < kNumSMs / 2 ? max_clusters : kNumSMs / 2;
    if (getenv("GFX_VERBOSE"))
      fprintf(stderr, "libgfx: K2 pair kernel: %d resident CTA pairs on device %d\n",
              resident[device], device);
  }
  const int cap = device >= 0 && device < 64 ? resident[device] : kNumSMs / 2;
  const int clusters = int(pairs < cap ? pairs : cap);
  kernel<<<2 * clusters, v7k::kWarps * 32, v7k::Smem::total, st>>>(maps, c, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx
