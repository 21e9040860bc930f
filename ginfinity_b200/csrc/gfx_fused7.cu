// One GINE layer in one kernel on CTA pairs, BANDED producers (seventh fused version).
//
//   h_out = h + LayerNorm(W2 relu(W1' z + b1') + b2),
//   z_i   = (1+eps) h_i + sum_{e: dst=i} relu(h[src_e] + table[type_e])
//
// Same skeleton as gfx_fused6.cu (CTA pairs, tcgen05 cta_group::2, three resident h-tile buffers
// that serve as neighbour source, residual and output staging, two z stages, epilogues in TMEM);
// what changes is how z is produced.  gfx_fused6.cu is bound by shared-memory bandwidth: per node
// its quarter-warp producers read 6.5 neighbour/self rows and 5.5 table rows of 256 B.  RNA graphs
// are BANDED: the incoming edges of nucleotide i are, in the reference's edge order,
//     (i-1, backbone fwd) (i+1, backbone rev) [(partner, pair fwd|rev)] (i-2, skip fwd) (i+2, skip rev)
// with some of them missing at molecule ends.  gfx_row_describe() classifies every CSR row once
// per chunk into a 32-bit descriptor (presence bits, pair type, partner index; or GENERIC when the
// row is anything else).  A producer WARP then owns a run of consecutive rows, each lane 4
// channels (8 bytes of a row): the run's rows i-2 .. i+2 are loaded ONCE into a register window
// (statically indexed, fully unrolled), the six table rows live in registers, only the pairing
// partner's row is fetched per node.  Shared-memory traffic of the producers drops from ~3.3 KB to
// ~0.9 KB per node, the CSR arrays are not read at all on the fast path (4 B instead of 29 B per
// node and layer), and the sum is taken in CSR order with the same arithmetic as gfx_fused6.cu, so
// the two kernels agree BIT FOR BIT.  GENERIC rows (sliced windows with context nodes, arbitrary
// graphs) are recomputed from the CSR arrays by a slow warp-per-row loop after the run.
//
// Warps per CTA (24; 80 registers per thread at launch = 61,440 for the CTA, re-balanced with
// setmaxnreg WITHIN that allocation -- a sum of the per-role limits above it deadlocks in
// setmaxnreg.inc -- epilogue A 64, epilogue B 80, producers 104, utility warpgroup 40):
//    0-3   epilogue A   D1 -> + b1, ReLU, fp16 -> A2 (in place)
//    4-11  epilogue B   two groups of 4, alternating tiles: D2 -> + b2, LayerNorm, + residual,
//                       in place in the h-tile buffer (the arithmetic of gfx_fused6.cu)
//   12-19  producers    warp w owns tile rows [16 w, 16 w + 16), two runs of 8 rows
//   20     MMA issuer (rank 0 issues for the pair; GEMM 1 as 8 MMAs of N = 256), 21 h-tile loader,
//   22     output store (both TMA)
#include <cstdlib>
#include <type_traits>

#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_pair.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v7 {

constexpr int HID = kMlpHidden, H = HID / 2;
constexpr int kTileM = 128;
constexpr int kKbBytes = kTileM * 128;        // one K block of a tile: [128 x 64] fp16
constexpr int kTileBytes = 2 * kKbBytes;      // a whole [128 x 128] fp16 tile
constexpr int kStages = 2, kHBufs = 3;
constexpr int kWPiece = 64 * 128;             // 64 weight rows x 64 columns (one CTA's share)
constexpr uint32_t kTmemCols = 512, kD2Col = 256;
constexpr int kEpiBWarp0 = 4, kProdWarp0 = 12, kProdWarps = 8, kMmaWarp = 20, kLoadWarp = 21,
              kStoreWarp = 22, kWarps = 24;
constexpr int kRowsPerWarp = kTileM / kProdWarps;     // 16 consecutive rows of a tile per producer warp
constexpr int kRun = 8;                               // rows per register window
static_assert(kRowsPerWarp % kRun == 0 && kRun % 4 == 0, "runs of whole partner groups");
// row descriptor (gfx_row_describe)
constexpr uint32_t kDescPrev = 1u, kDescNext = 2u, kDescPair = 4u, kDescPairRev = 8u, kDescPrev2 = 16u,
                   kDescNext2 = 32u, kDescGeneric = 0x80000000u;
constexpr int kDescPartnerShift = 6, kDescPartnerBits = 25;
constexpr uint32_t kDescPartnerMask = (1u << kDescPartnerBits) - 1u;

enum Bar {
  kBarWLocal = 0, kBarWReady = 1, kBarHFull = 2, kBarHEmpty = 5, kBarOReady = 8, kBarA1Full = 11,
  kBarA1Empty = 13, kBarD1aFull = 15, kBarD1bFull = 16, kBarA2aFull = 17, kBarA2bFull = 18,
  kBarD2Full = 19, kBarD2Empty = 21, kNumBars = 23
};

struct Smem {
  static constexpr int off_w1 = 0;                                   // [kb 2] x 16 KB (128 rows)
  static constexpr int off_w2 = off_w1 + 4 * kWPiece;                // [kb 4] x 8 KB
  static constexpr int off_z = off_w2 + 4 * kWPiece;                 // 2 stages x 32 KB
  static constexpr int off_h = off_z + kStages * kTileBytes;         // 3 tiles x 32 KB
  static constexpr int off_bar = off_h + kHBufs * kTileBytes;
  static constexpr int off_tmem = off_bar + kNumBars * 8;
  static constexpr int total = off_tmem + 8;
};
static_assert(Smem::total <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct alignas(64) Maps {
  CUtensorMap h, out;        // [n, 128] fp16, box 64 x 128, SWIZZLE_128B
};

struct Consts {
  float b1[HID];
};

struct Args {
  const __half *h;
  const int32_t *row_ptr, *col_src;
  const uint8_t *col_type;
  const uint32_t *desc;      // [n] row descriptors
  uint32_t sleep_ns;         // sleep between barrier polls of the warp-wide waits
  const __half *table16, *w1_img, *w2_img;
  const float *b2, *g, *b;   // device vectors of this layer
  int64_t n;
  int edge_dim;
  float eps1;
  long long *trace;          // developer timeline (tools/fused_trace.py); null in production
};

// parked wait (suspend-time hint) on a local barrier
__device__ __forceinline__ void mbar_wait_c(uint64_t *bar, uint32_t parity) {
  mbar_wait_parked(bar, parity);
}
// Wait of a whole warp (or of a utility thread off the critical path) with a plain sleep between
// polls.  The parked form (try_wait with a suspend-time hint) compiles to TRYWAIT + NANOSLEEP.SYNCS,
// and that sleep ends on ANY barrier activity of the CTA: in this kernel a waiting warp re-polled
// every ~50 cycles, and those loops were 30 % of all issued instructions of an issue-bound kernel
// (ncu source page of the first banded version).
__device__ __forceinline__ void mbar_wait_s(uint64_t *bar, uint32_t parity, uint32_t sleep_ns) {
  while (!mbar_try_wait(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);
  }
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

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
__device__ __forceinline__ uint32_t hfma2_relu_add(uint32_t x, uint32_t t) {
  uint32_t r;
  asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0x3c003c00u), "r"(t));
  return r;
}
__device__ __forceinline__ void add_pair(float &a0, float &a1, uint32_t m) {
  asm volatile("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.f16 %0, lo, %0;\n"
      "add.rn.f32.f16 %1, hi, %1;\n}"
      : "+f"(a0), "+f"(a1)
      : "r"(m));
}
// one message on this lane's 4 channels: fp16 relu(x + t), summed in fp32
__device__ __forceinline__ void add_message4(float *acc, const uint2 &nb, const uint2 &tb) {
  add_pair(acc[0], acc[1], hfma2_relu_add(nb.x, tb.x));
  add_pair(acc[2], acc[3], hfma2_relu_add(nb.y, tb.y));
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4 &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t saddr, const uint2 &v) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(v.x), "r"(v.y) : "memory");
}
// 8 bytes of a row: from the resident h tile when the row lies in this tile, else from global
// memory (L2); one predicated instruction of each kind writing the same registers
__device__ __forceinline__ uint2 ld_tile_or_global8(uint32_t in_tile, uint32_t saddr, const uint2 *gptr) {
  uint2 v;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %2, 0;\n"
      "@q ld.shared.v2.u32 {%0, %1}, [%3];\n"
      "@!q ld.global.nc.v2.u32 {%0, %1}, [%4];\n"
      "}\n"
      : "=r"(v.x), "=r"(v.y)
      : "r"(in_tile), "r"(saddr), "l"(gptr));
  return v;
}
// byte offset of 16-byte chunk `c8` (0..7) of row `r` inside one swizzled K block
__device__ __forceinline__ uint32_t sw_off(int r, int c8) {
  return uint32_t(r) * 128u + (uint32_t((c8 ^ r) & 7) << 4);
}

// D1[:, HALF*128 .. +128) -> bias + ReLU -> fp16 -> A2[:, HALF*64 .. +64), 32 columns at a time
template <int HALF>
__device__ __forceinline__ void epi_a(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      int lane, uint32_t leader_bar, uint32_t sleep_ns) {
  constexpr int col0 = HALF * H;
  mbar_wait_s(bar + (HALF ? kBarD1bFull : kBarD1aFull), ph, sleep_ns);
  tc_fence_after();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[32];
    tmem_ld32(trow + col0 + 32 * q, v);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      pk[j] = relu_pack2(v[2 * j] + c.b1[col0 + 32 * q + 2 * j],
                         v[2 * j + 1] + c.b1[col0 + 32 * q + 2 * j + 1]);
    tmem_st16(trow + col0 / 2 + 16 * q, pk);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(leader_bar);
}

// developer timeline: event `ev` of tile iteration `it`, CTA 0 only
__device__ __forceinline__ void trace_ev(const Args &p, uint32_t it, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && it < 64) p.trace[it * 16 + ev] = clock64();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWarps * 32, 1)
fused_banded_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Consts c, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2, *zs = smem + L::off_z;
  uint8_t *hs = smem + L::off_h;
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
    for (int s = 0; s < kHBufs; ++s) {
      mbar_init(bar + kBarHFull + s, 1);
      mbar_init(bar + kBarHEmpty + s, 1);
      mbar_init(bar + kBarOReady + s, 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar + kBarA1Full + s, 2 * kProdWarps);
      mbar_init(bar + kBarA1Empty + s, 1);
      mbar_init(bar + kBarD2Full + s, 1);
      mbar_init(bar + kBarD2Empty + s, 8);
    }
    mbar_init(bar + kBarD1aFull, 1);
    mbar_init(bar + kBarD1bFull, 1);
    mbar_init(bar + kBarA2aFull, 8);
    mbar_init(bar + kBarA2bFull, 8);
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
  // barriers of rank 0 that collect arrivals from both CTAs
  auto leader = [&](int b) { return map_to_cta(smem_u32(bar + b), 0); };

  if (warp < kEpiBWarp0) {
    // ================= epilogue A =================================================
    reg_dec<64>();
    const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
    const uint32_t a2a = leader(kBarA2aFull), a2b = leader(kBarA2bFull);
    uint32_t it = 0;
    for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
      if (tid == 0) trace_ev(p, it, 6);
      epi_a<0>(c, trow, bar, it & 1, lane, a2a, p.sleep_ns);
      epi_a<1>(c, trow, bar, it & 1, lane, a2b, p.sleep_ns);
      if (tid == 0) trace_ev(p, it, 7);
    }
  } else if (warp < kProdWarp0) {
    // ================= epilogue B (group g takes iterations it % 2 == g) ===========
    const int quad = warp & 3, g = (warp - kEpiBWarp0) >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16) + kD2Col + uint32_t(g) * kHidden;
    const uint32_t d2e = leader(kBarD2Empty + g);
    const float4 *b2v = reinterpret_cast<const float4 *>(p.b2);
    const float4 *gv = reinterpret_cast<const float4 *>(p.g);
    const float4 *bv = reinterpret_cast<const float4 *>(p.b);
    const int r = quad * 32 + lane;
    uint32_t it = 0;
    for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
      if (int(it & 1) != g) continue;
      const uint32_t hb = it % kHBufs;
      mbar_wait_s(bar + kBarD2Full + g, (it >> 1) & 1, p.sleep_ns);
      tc_fence_after();
      if (lane == 0 && quad == 0) trace_ev(p, it, 8);
      // 16 columns at a time, loops NOT unrolled: an unrolled body lets the compiler hoist the
      // vector loads of several chunks and spill (64 registers per thread)
      float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll 1
      for (int q = 0; q < 8; ++q) {
        float u[16];
        tmem_ld16(trow + 16 * q, u);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 bb = __ldg(b2v + 4 * q + j4);
          const float t0 = u[4 * j4] + bb.x, t1 = u[4 * j4 + 1] + bb.y;
          const float t2 = u[4 * j4 + 2] + bb.z, t3 = u[4 * j4 + 3] + bb.w;
          s1[0] += t0; s1[1] += t1; s1[0] += t2; s1[1] += t3;
          s2[0] = fmaf(t0, t0, s2[0]); s2[1] = fmaf(t1, t1, s2[1]);
          s2[0] = fmaf(t2, t2, s2[0]); s2[1] = fmaf(t3, t3, s2[1]);
        }
      }
      const float mean = (s1[0] + s1[1]) * (1.f / kHidden);
      const float var = fmaxf((s2[0] + s2[1]) * (1.f / kHidden) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      const float nm = -mean * rstd;
      mbar_wait_s(bar + kBarHFull + hb, (it / kHBufs) & 1, 0);   // long complete: visibility only
      const uint32_t hrow = smem_u32(hs) + hb * kTileBytes;
#pragma unroll 1
      for (int q = 0; q < 8; ++q) {
        float u[16];
        tmem_ld16(trow + 16 * q, u);
        tmem_ld_wait();
        if (q == 7) {                                     // D2 fully read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(d2e);
        }
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
          const int c16 = q * 2 + gi;                     // 16-byte chunk of the 256-byte row
          const uint32_t cell = hrow + uint32_t(c16 >> 3) * kKbBytes + sw_off(r, c16 & 7);
          const uint4 raw = lds128(cell);
          const __half2 *hp = reinterpret_cast<const __half2 *>(&raw);
          float o[8];
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int j = gi * 8 + 4 * w, c4 = 4 * q + 2 * gi + w;
            const float4 bb = __ldg(b2v + c4), gg = __ldg(gv + c4), be = __ldg(bv + c4);
            const float2 ra = __half22float2(hp[2 * w]), rb = __half22float2(hp[2 * w + 1]);
            o[4 * w] = fmaf(fmaf(u[j] + bb.x, rstd, nm), gg.x, ra.x + be.x);
            o[4 * w + 1] = fmaf(fmaf(u[j + 1] + bb.y, rstd, nm), gg.y, ra.y + be.y);
            o[4 * w + 2] = fmaf(fmaf(u[j + 2] + bb.z, rstd, nm), gg.z, rb.x + be.z);
            o[4 * w + 3] = fmaf(fmaf(u[j + 3] + bb.w, rstd, nm), gg.w, rb.y + be.w);
          }
          sts128(cell, make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]),
                                  pack2(o[6], o[7])));
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + kBarOReady + hb);
      if (lane == 0 && quad == 0) trace_ev(p, it, 9);
    }
  } else if (warp < kMmaWarp) {
    // ================= producers: banded aggregation into the z stage ================
    reg_inc<104>();
    const int pw = warp - kProdWarp0;                 // rows [16 pw, 16 pw + 16) of every tile
    // this lane's 4 channels = 8 bytes of a row: 16-byte chunk lane / 2 of the 256-byte row
    const uint32_t kboff = uint32_t(lane >> 4) * kKbBytes;
    const uint32_t c8 = uint32_t(lane >> 1) & 7u, odd8 = uint32_t(lane & 1) * 8u;
    auto cell = [&](int r) -> uint32_t {              // byte offset of this lane's piece of tile row r
      return kboff + uint32_t(r) * 128u + (((c8 ^ uint32_t(r)) & 7u) << 4) + odd8;
    };
    const uint2 *hg = reinterpret_cast<const uint2 *>(p.h) + lane;     // row r -> hg[r * 32]
    const uint32_t a1f[2] = {leader(kBarA1Full), leader(kBarA1Full + 1)};
    const uint2 kNone = make_uint2(0xFBFFFBFFu, 0xFBFFFBFFu);          // -65504: relu(x + it) = +0
    uint2 tb[6];                                                       // table rows of types 0..5
#pragma unroll
    for (int k = 0; k < 6; ++k)
      tb[k] = reinterpret_cast<const uint2 *>(p.table16 + k * kHidden)[lane];

    // lane l < 16 holds the descriptor of the warp's row l, fetched one tile ahead
    auto fetch_desc = [&](int pr) -> uint32_t {
      if (pr >= pairs || lane >= kRowsPerWarp) return 0u;
      const int row = (2 * pr + int(rank)) * kTileM + kRowsPerWarp * pw + lane;
      return row < n ? p.desc[row] : 0u;
    };
    // the two rows just outside the tile that the first / last warp's windows reach (global memory)
    auto fetch_halo = [&](int pr, uint2 &x0, uint2 &x1) {
      x0 = x1 = make_uint2(0u, 0u);
      if (pr >= pairs || (pw != 0 && pw != kProdWarps - 1)) return;
      const int r0 = (2 * pr + int(rank)) * kTileM;
      const int g0 = pw == 0 ? r0 - 2 : r0 + kTileM, g1 = g0 + 1;
      if (g0 >= 0 && g0 < n) x0 = __ldg(hg + int64_t(g0) * 32);
      if (g1 >= 0 && g1 < n) x1 = __ldg(hg + int64_t(g1) * 32);
    };
    uint32_t dnext = fetch_desc(cluster_id);
    uint2 hx0, hx1;
    fetch_halo(cluster_id, hx0, hx1);
    uint32_t it = 0;
    for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
      const uint32_t s = it & 1, hb = it % kHBufs;
      const int row0 = (2 * pair + int(rank)) * kTileM;
      const uint32_t d = dnext;
      const uint2 halo0 = hx0, halo1 = hx1;
      dnext = fetch_desc(pair + clusters);
      fetch_halo(pair + clusters, hx0, hx1);
      const uint32_t hbase = smem_u32(hs) + hb * kTileBytes;
      const uint32_t zbase = smem_u32(zs) + s * kTileBytes;
      mbar_wait_s(bar + kBarHFull + hb, (it / kHBufs) & 1, p.sleep_ns);
      mbar_wait_s(bar + kBarA1Empty + s, ((it >> 1) & 1) ^ 1, p.sleep_ns);
      if (warp == kProdWarp0 && lane == 0) trace_ev(p, it, 0);
      // partner row of a row with descriptor dj (a row without a pair reads itself; its message is +0)
      auto partner = [&](uint32_t dj, int idx) -> uint2 {
        const int self = kRowsPerWarp * pw + idx;
        const int src = (dj & kDescPair) ? int((dj >> kDescPartnerShift) & kDescPartnerMask) : row0 + self;
        const uint32_t local = uint32_t(src - row0);
        const uint32_t in_tile = local < uint32_t(kTileM) ? 1u : 0u;
        return ld_tile_or_global8(in_tile, hbase + cell(int(local & (kTileM - 1))), hg + int64_t(src) * 32);
      };
#pragma unroll 1
      for (int run = 0; run < kRowsPerWarp / kRun; ++run) {
        const int base = kRowsPerWarp * pw + kRun * run;   // first tile row of the run
        uint32_t dj[kRun];                                 // the run's descriptors, warp-uniform
#pragma unroll
        for (int j = 0; j < kRun; ++j) dj[j] = __shfl_sync(0xffffffffu, d, kRun * run + j);
        uint2 pr[2][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) pr[0][q] = partner(dj[q], kRun * run + q);
        // window: tile rows base - 2 .. base + kRun + 1.  base is a multiple of 8, so the swizzle
        // term of row base - 2 + k depends on k only: one XOR with an immediate per row.  Only
        // the first two / last two rows of the window can lie outside the tile (first / last run
        // of the tile: warp-uniform), where the prefetched halo rows stand in.
        static_assert(kRun == 8, "the swizzle pattern repeats every 8 rows");
        const uint32_t wbase = hbase + kboff + odd8 + uint32_t(base) * 128u;   // row `base`, chunk 0
        const uint32_t c8s = c8 << 4;
        auto wrow = [&](int k) -> uint2 {                   // k = 0 .. kRun + 3 <-> row base - 2 + k
          return lds64(wbase + uint32_t((k - 2) * 128) + (c8s ^ (uint32_t((k - 2) & 7) << 4)));
        };
        const bool first = base == 0, last = base + kRun == kTileM;
        uint2 w[kRun + 4];
        w[0] = halo0;
        w[1] = halo1;
        if (!first) {
          w[0] = wrow(0);
          w[1] = wrow(1);
        }
#pragma unroll
        for (int k = 2; k < kRun + 2; ++k) w[k] = wrow(k);
        w[kRun + 2] = halo0;
        w[kRun + 3] = halo1;
        if (!last) {
          w[kRun + 2] = wrow(kRun + 2);
          w[kRun + 3] = wrow(kRun + 3);
        }
        // INTERIOR: every row of the run has its four backbone / skip neighbours (all but the runs
        // at molecule ends): their table rows are used as they are, no selects
        auto nodes = [&](auto interior_c) {
          constexpr bool INTERIOR = decltype(interior_c)::value;
#pragma unroll
          for (int j = 0; j < kRun; ++j) {
            if ((j & 3) == 0 && j + 4 < kRun) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                pr[((j >> 2) + 1) & 1][q] = partner(dj[j + 4 + q], kRun * run + j + 4 + q);
            }
            const uint32_t dd = dj[j];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            uint2 tt;
            tt = INTERIOR || (dd & kDescPrev) ? tb[0] : kNone;
            add_message4(acc, w[j + 1], tt);
            tt = INTERIOR || (dd & kDescNext) ? tb[1] : kNone;
            add_message4(acc, w[j + 3], tt);
            tt = (dd & kDescPairRev) ? tb[3] : tb[2];
            tt = (dd & kDescPair) ? tt : kNone;
            add_message4(acc, pr[(j >> 2) & 1][j & 3], tt);
            tt = INTERIOR || (dd & kDescPrev2) ? tb[4] : kNone;
            add_message4(acc, w[j], tt);
            tt = INTERIOR || (dd & kDescNext2) ? tb[5] : kNone;
            add_message4(acc, w[j + 4], tt);
            const __half2 *sv = reinterpret_cast<const __half2 *>(&w[j + 2]);
            const float2 f0 = __half22float2(sv[0]), f1 = __half22float2(sv[1]);
            uint2 o;
            o.x = pack2(fmaf(p.eps1, f0.x, acc[0]), fmaf(p.eps1, f0.y, acc[1]));
            o.y = pack2(fmaf(p.eps1, f1.x, acc[2]), fmaf(p.eps1, f1.y, acc[3]));
            sts64(zbase + kboff + odd8 + uint32_t(base + j) * 128u + (c8s ^ (uint32_t(j & 7) << 4)), o);
          }
        };
        uint32_t all = dj[0];
#pragma unroll
        for (int j = 1; j < kRun; ++j) all &= dj[j];
        constexpr uint32_t kBackbone = kDescPrev | kDescNext | kDescPrev2 | kDescNext2;
        if ((all & kBackbone) == kBackbone)
          nodes(std::true_type{});
        else
          nodes(std::false_type{});
      }
      // rows that are not banded: recomputed from the CSR arrays, one row at a time (warp-uniform)
      if (__ballot_sync(0xffffffffu, (d & kDescGeneric) != 0u) != 0u) {
#pragma unroll 1
        for (int idx = 0; idx < kRowsPerWarp; ++idx) {
          if (!(__shfl_sync(0xffffffffu, d, idx) & kDescGeneric)) continue;
          const int lr = kRowsPerWarp * pw + idx, row = row0 + lr;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          const int beg = p.row_ptr[row], end = p.row_ptr[row + 1];
          for (int e = beg; e < end; ++e) {
            const int src = p.col_src[e];
            const uint2 tt = reinterpret_cast<const uint2 *>(p.table16 + int(p.col_type[e]) * kHidden)[lane];
            const uint32_t local = uint32_t(src - row0);
            const uint2 v = local < uint32_t(kTileM) ? lds64(hbase + cell(int(local))) : __ldg(hg + int64_t(src) * 32);
            add_message4(acc, v, tt);
          }
          const uint2 self = lds64(hbase + cell(lr));
          const __half2 *sv = reinterpret_cast<const __half2 *>(&self);
          const float2 f0 = __half22float2(sv[0]), f1 = __half22float2(sv[1]);
          uint2 o;
          o.x = pack2(fmaf(p.eps1, f0.x, acc[0]), fmaf(p.eps1, f0.y, acc[1]));
          o.y = pack2(fmaf(p.eps1, f1.x, acc[2]), fmaf(p.eps1, f1.y, acc[3]));
          sts64(zbase + cell(lr), o);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(a1f[s]);
      if (warp == kProdWarp0 && lane == 0) trace_ev(p, it, 1);
    }
  } else if (warp == kMmaWarp) {
    // ================= weights (both CTAs) + MMA issue (rank 0) ======================
    reg_dec<40>();
    if (lane == 0) {
      mbar_arrive_expect_tx(bar + kBarWLocal, 8 * kWPiece);
      const uint8_t *w1g = reinterpret_cast<const uint8_t *>(p.w1_img);
      const uint8_t *w2g = reinterpret_cast<const uint8_t *>(p.w2_img);
      // Each CTA supplies W1' rows [128 rank, +128) of every K block (two 8 KB pieces, contiguous:
      // one B operand of an N = 256 MMA).  In pair mode an MMA costs ~128 cycles whether N is 128
      // or 256 (measured with gfx_umma7.cu), so GEMM 1 runs as 8 MMAs of N = 256, not 16 of N = 128.
      for (int kb = 0; kb < 2; ++kb)
        for (int half = 0; half < 2; ++half)
          bulk_g2s(w1s + (kb * 2 + half) * kWPiece,
                   w1g + kb * (HID * 128) +
                       (H * int(rank) + 64 * half) * 128,
                   kWPiece, bar + kBarWLocal);
      for (int kb = 0; kb < 4; ++kb)              // rows [64 rank, +64) of K block kb
        bulk_g2s(w2s + kb * kWPiece, w2g + kb * (kHidden * 128) + 64 * int(rank) * 128, kWPiece,
                 bar + kBarWLocal);
      mbar_wait_parked(bar + kBarWLocal, 0);
      mbar_arrive_cluster(leader(kBarWReady));
      if (rank == 0) {
        mbar_wait_c(bar + kBarWReady, 0);
        constexpr uint32_t idesc = idesc_f16(2 * kTileM, H);     // M = 256 over the pair, N = 128
        const uint32_t w1a = smem_u32(w1s), w2a = smem_u32(w2s);
        uint32_t it = 0;
        for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
          const uint32_t s = it & 1, g = it & 1, ph2 = (it >> 1) & 1, ph = it & 1;
          const uint32_t za = smem_u32(zs) + s * kTileBytes;
          mbar_wait_c(bar + kBarA1Full + s, ph2);
          tc_fence_after();
          trace_ev(p, it, 2);
          constexpr uint32_t idesc_w = idesc_f16(2 * kTileM, HID);   // GEMM 1: N = 256
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int kb = kk >> 2, k = kk & 3;
            const uint64_t da = smem_desc_sw128(za + kb * kKbBytes + k * 32);
            const uint64_t db = smem_desc_sw128(w1a + kb * 2 * kWPiece + k * 32);
            mma2_f16_ss(tmem, da, db, idesc_w, kk != 0);
          }
          mma2_commit(bar + kBarD1aFull);
          mma2_commit(bar + kBarD1bFull);
          mma2_commit(bar + kBarA1Empty + s);          // z consumed in both CTAs
          trace_ev(p, it, 3);
          mbar_wait_c(bar + kBarA2aFull, ph);
          mbar_wait_c(bar + kBarD2Empty + g, ph2 ^ 1);
          tc_fence_after();
          trace_ev(p, it, 4);
#pragma unroll
          for (int kk = 0; kk < HID / 16; ++kk) {
            if (kk == H / 16) {
              mbar_wait_c(bar + kBarA2bFull, ph);
              tc_fence_after();
            }
            const uint64_t db = smem_desc_sw128(w2a + (kk >> 2) * kWPiece + (kk & 3) * 32);
            mma2_f16_ts(tmem + kD2Col + g * kHidden, tmem + kk * 8, db, idesc, kk != 0);
          }
          mma2_commit(bar + kBarD2Full + g);
          trace_ev(p, it, 5);
        }
      }
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    // ================= h tiles (TMA) =================================================
    reg_dec<40>();
    if (lane == 0) {
      prefetch_tmap(&maps.h);
      uint32_t it = 0;
      for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
        const uint32_t hb = it % kHBufs;
        const int row0 = (2 * pair + int(rank)) * kTileM;
        uint8_t *dst = hs + hb * kTileBytes;
        mbar_wait_s(bar + kBarHEmpty + hb, ((it / kHBufs) & 1) ^ 1, p.sleep_ns);
        trace_ev(p, it, 12);
        mbar_arrive_expect_tx(bar + kBarHFull + hb, kTileBytes);
        tma_load_2d(dst, &maps.h, 0, row0, bar + kBarHFull + hb);
        tma_load_2d(dst + kKbBytes, &maps.h, 64, row0, bar + kBarHFull + hb);
      }
    }
    __syncwarp();
  } else if (warp == kStoreWarp) {
    // ================= output store (TMA) ============================================
    reg_dec<40>();
    if (lane == 0) {
      prefetch_tmap(&maps.out);
      uint32_t it = 0;
      for (int pair = cluster_id; pair < pairs; pair += clusters, ++it) {
        const uint32_t hb = it % kHBufs;
        const int row0 = (2 * pair + int(rank)) * kTileM;
        const uint8_t *src = hs + hb * kTileBytes;
        mbar_wait_s(bar + kBarOReady + hb, (it / kHBufs) & 1, p.sleep_ns);
        trace_ev(p, it, 10);
        if (row0 < n) {
          tma_store_2d(&maps.out, 0, row0, src);
          tma_store_2d(&maps.out, 64, row0, src + kKbBytes);
        }
        bulk_commit();
        bulk_wait_read<0>();                      // shared memory has been read: the buffer is free
        mbar_arrive(bar + kBarHEmpty + hb);
        trace_ev(p, it, 11);
      }
      bulk_wait_all();
    }
    __syncwarp();
  } else {
    reg_dec<40>();                                // spare warp of the utility warpgroup
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // the peer may still be arriving on our barriers
  if (warp == kMmaWarp) tmem_dealloc2(tmem, kTmemCols);
  if (p.trace != nullptr && tid == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.trace[1024 + 2 * blockIdx.x + 1] = (long long)ns;
  }
}

}  // namespace v7

// ---- row descriptors ---------------------------------------------------------------------------
// One thread per CSR row: does the row read, in order, (i-1, type 0) (i+1, type 1)
// [(any, type 2|3)] (i-2, type 4) (i+2, type 5), each optional, and nothing else?
__global__ void __launch_bounds__(256)
row_describe_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_src,
                    const uint8_t *__restrict__ col_type, int64_t n, uint32_t *__restrict__ desc) {
  using namespace v7;
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int end = row_ptr[i + 1];
  int k = row_ptr[i];
  uint32_t d = 0;
  auto is = [&](int64_t src, int type) { return k < end && col_src[k] == src && col_type[k] == type; };
  if (is(i - 1, 0)) { d |= kDescPrev; ++k; }
  if (is(i + 1, 1)) { d |= kDescNext; ++k; }
  if (k < end && (col_type[k] == 2 || col_type[k] == 3) && col_src[k] >= 0 &&
      uint32_t(col_src[k]) <= kDescPartnerMask) {
    d |= kDescPair | (col_type[k] == 3 ? kDescPairRev : 0u) | (uint32_t(col_src[k]) << kDescPartnerShift);
    ++k;
  }
  if (is(i - 2, 4)) { d |= kDescPrev2; ++k; }
  if (is(i + 2, 5)) { d |= kDescNext2; ++k; }
  desc[i] = k == end ? d : kDescGeneric;
}

extern long long *g_trace;              // gfx_fused6.cu (gfx_debug_fused_trace)

int fused7_layer(const gfx_model *m, int layer, const __half *h, const int32_t *row_ptr,
                 const int32_t *col_src, const uint8_t *col_type, const uint32_t *desc, int64_t n,
                 __half *h_out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(h_out)) & 15)
    return fail(GFX_ERR_ARGUMENT, "fused layer: activation buffers must be 16-byte aligned");
  if (h == h_out) return fail(GFX_ERR_ARGUMENT, "fused layer: h and h_out must not alias");
  if (n > (int64_t(1) << v7::kDescPartnerBits))
    return fail(GFX_ERR_UNSUPPORTED, "fused layer (banded): at most 2^25 nodes per call");
  if (m->edge_dim < 6)
    return fail(GFX_ERR_UNSUPPORTED, "fused layer (banded): needs the six backbone / pair / skip edge types");
  v7::Maps maps;
  int rc = tma::make_rows128_map(&maps.h, h, n, v7::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.out, h_out, n, v7::kTileM);
  if (rc) return rc;
  v7::Consts c;
  const gfx_host_vectors &hv = m->host;
  for (int i = 0; i < kMlpHidden; ++i) c.b1[i] = hv.b1[size_t(layer) * kMlpHidden + i];
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  v7::Args a{};
  a.h = h; a.row_ptr = row_ptr; a.col_src = col_src; a.col_type = col_type; a.desc = desc;
  a.table16 = m->table16 + size_t(layer) * m->edge_dim * kHidden;
  a.w1_img = m->w1_img + wi; a.w2_img = m->w2_img + wi;
  a.b2 = m->b2 + size_t(layer) * kHidden;
  a.g = m->ln_g + size_t(layer) * kHidden;
  a.b = m->ln_b + size_t(layer) * kHidden;
  a.n = n; a.edge_dim = m->edge_dim; a.eps1 = m->eps1[layer];
  a.trace = g_trace;
  static const uint32_t sleep_ns = [] {
    const char *v = getenv("GFX_FUSED_SLEEP_NS");   // developer switch
    return uint32_t(v ? atoi(v) : 64);
  }();
  a.sleep_ns = sleep_ns;
  auto kernel = v7::fused_banded_kernel;
  GFX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v7::Smem::total));
  const int64_t tiles = (n + v7::kTileM - 1) / v7::kTileM;
  const int64_t pairs = (tiles + 1) / 2;
  // as many CTA pairs as this GPU can hold at once (see gfx_fused6.cu)
  static int resident[64] = {};                   // per device; 0 = not asked yet
  int device = 0;
  GFX_CUDA(cudaGetDevice(&device));
  if (device >= 0 && device < 64 && resident[device] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs, 1, 1);
    cfg.blockDim = dim3(v7::kWarps * 32, 1, 1);
    cfg.dynamicSmemBytes = v7::Smem::total;
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
  }
  const int cap = device >= 0 && device < 64 ? resident[device] : kNumSMs / 2;
  const int clusters = int(pairs < cap ? pairs : cap);
  kernel<<<2 * clusters, v7::kWarps * 32, v7::Smem::total, st>>>(maps, c, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx

extern "C" int gfx_row_describe(const int32_t *row_ptr, const int32_t *col_src,
                                const uint8_t *col_type, int64_t n, uint32_t *desc, void *stream) {
  using namespace gfx;
  if (n <= 0) return GFX_OK;
  if (!row_ptr || !desc) return fail(GFX_ERR_ARGUMENT, "gfx_row_describe: null array");
  cudaStream_t st = as_stream(stream);
  StageScope scope(GFX_STAGE_CSR, st, 1);
  row_describe_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(row_ptr, col_src, col_type, n, desc);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

extern "C" int gfx_layer_fused_banded(const gfx_model *m, int layer, const void *h,
                                      const int32_t *row_ptr, const int32_t *col_src,
                                      const uint8_t *col_type, const uint32_t *desc, int64_t n,
                                      void *h_out, void *stream) {
  using namespace gfx;
  if (!m || layer < 0 || layer >= m->layers)
    return fail(GFX_ERR_ARGUMENT, "gfx_layer_fused_banded: bad model or layer");
  if (n <= 0) return GFX_OK;
  if (!desc) return fail(GFX_ERR_ARGUMENT, "gfx_layer_fused_banded: null row descriptors");
  cudaStream_t st = as_stream(stream);
  StageScope scope(GFX_STAGE_FUSED_LAYER, st, 1);
  return fused7_layer(m, layer, static_cast<const __half *>(h), row_ptr, col_src, col_type, desc, n,
                      static_cast<__half *>(h_out), st);
}
