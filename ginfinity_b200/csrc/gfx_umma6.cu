// K2 (MLP + LayerNorm + residual), sixth tcgen05 version: three stage buffers, each cycling
// z tile -> residual tile -> output tile.
//
// Timeline of gfx_umma4.cu (tools/k2_trace.py, cycles per 128-row tile): GEMM 1 issues in 1.4 k,
// GEMM 2 in 0.95 k, but the period is 4.45 k, and it is the same whether the working set is in L2
// or in HBM (tools/k2_probe.py) -- the kernel is bound by a latency chain, not by bandwidth.  The
// chain runs through its single residual/output tile R: store of tile t reads R -> TMA load of
// tile t+1's residual (an HBM round trip, ~2.5 k) -> epilogue B of t+1 can finish (1.5 k) -> store.
// Here R is gone: a stage buffer takes its tile's z, then -- as soon as GEMM 1 has consumed it --
// the tile's residual rows, in which epilogue B writes the output in place, then the TMA store reads
// it.  One buffer's cycle is ~8.5 k cycles (two HBM round trips), so with THREE of them (the 32 KB
// of R buy the third) the period they allow is 2.8 k, below what the tensor pipe and the
// epilogues need.  (gfx_fused5.cu has this life cycle with two buffers: 4.2 k.)
//
// Warps: 0-7 epilogue A, 8-15 epilogue B, 16 MMA issuer, 17 z loader, 18 residual loader,
// 19 output store (all TMA).  Everything else as in gfx_umma4.cu.
#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v6k {

constexpr int HID = kMlpHidden, H = HID / 2;
constexpr int kTileM = 128;
constexpr int kTileBytes = kTileM * 128;      // [128 x 64] fp16 box
constexpr int kA1Bytes = 2 * kTileBytes;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kA2Col = 256, kD2Col = 384;
constexpr int kEpiBWarp0 = 8, kMmaWarp = 16, kLoadWarp = 17, kResWarp = 18, kStoreWarp = 19, kWarps = 20;
constexpr int kStages = 3;

enum Bar {
  kBarW = 0, kBarD1aFull = 1, kBarD1bFull = 2, kBarA2aFull = 3, kBarA2bFull = 4, kBarD2Full = 5,
  kBarD2Empty = 6,
  kBarA1Full = 7,            // [stage 3]      z landed
  kBarA1Empty = 10,          // [stage 3]      GEMM 1 has consumed z
  kBarStageFree = 13,        // [stage 3]      the store has read the output
  kBarRFull = 16,            // [stage 3][2]   residual half landed
  kBarOReady = 22,           // [stage 3][2]   output half written
  kNumBars = 28
};

struct Smem {
  static constexpr int w_bytes = HID * kHidden * 2;                 // each weight image
  static constexpr int off_w1 = 0;
  static constexpr int off_w2 = off_w1 + w_bytes;
  static constexpr int off_a1 = off_w2 + w_bytes;                   // 3 stages x 32 KB
  static constexpr int off_xs = off_a1 + kStages * kA1Bytes;        // float2[2][128] partial sums
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
  if (p.trace != nullptr && blockIdx.x == 0 && it < 64) p.trace[it * 16 + ev] = clock64();
}

// {lo, hi} -> fp16x2 with ReLU folded into the conversion
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

// D1[:, HALF*128 + CH*64 .. +64) -> bias + ReLU -> fp16 -> A2
template <int HALF, int CH>
__device__ __forceinline__ void epi_a(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      int lane) {
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
  if (lane == 0) mbar_arrive(bar + (HALF ? kBarA2bFull : kBarA2aFull));
}

// D2[:, CH*64 .. +64) -> + b2 -> LayerNorm (stats shared with the other column
// half) -> * g + b + residual, in place in the residual tile R
template <int CH>
__device__ __forceinline__ void epi_b(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      uint32_t st, uint32_t sph, int lane, int quad, float2 *xs,
                                      uint8_t *rs) {   // rs: stage buffer `st` of this tile
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
  if (lane == 0) mbar_arrive(bar + kBarD2Empty);        // accumulator is in registers now
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
  uint8_t *rrow = rs + CH * kTileBytes + r * 128;
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

__global__ void __launch_bounds__(kWarps * 32, 1)
umma6_mlp_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Consts c, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2, *a1s = smem + L::off_a1;
  float2 *xs = reinterpret_cast<float2 *>(smem + L::off_xs);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::off_bar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
  } else if (tid == 0) {
    mbar_init(bar + kBarW, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar + kBarA1Full + s, 1);
      mbar_init(bar + kBarA1Empty + s, 1);
      mbar_init(bar + kBarStageFree + s, 1);
      for (int j = 0; j < 2; ++j) {
        mbar_init(bar + kBarRFull + 2 * s + j, 1);
        mbar_init(bar + kBarOReady + 2 * s + j, 4);
      }
    }
    mbar_init(bar + kBarD1aFull, 1);
    mbar_init(bar + kBarD1bFull, 1);
    mbar_init(bar + kBarA2aFull, 8);
    mbar_init(bar + kBarA2bFull, 8);
    mbar_init(bar + kBarD2Full, 1);
    mbar_init(bar + kBarD2Empty, 8);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t tiles = (p.n + kTileM - 1) / kTileM;

  if (warp < kEpiBWarp0) {
    // ================= epilogue A: D1 -> bias + ReLU -> fp16 -> A2 (TMEM) =========
    const int quad = warp & 3, ch = warp >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      if (ch == 0) {
        epi_a<0, 0>(c, trow, bar, ph, lane);
        if (tid == 0) trace_ev(p, it, 6);
        epi_a<1, 0>(c, trow, bar, ph, lane);
        if (tid == 0) trace_ev(p, it, 7);
      } else {
        epi_a<0, 1>(c, trow, bar, ph, lane);
        epi_a<1, 1>(c, trow, bar, ph, lane);
      }
    }
  } else if (warp < kMmaWarp) {
    // ===== epilogue B: D2 -> bias + LayerNorm, + residual from R, in place in R =====
    const int quad = warp & 3, ch = (warp - kEpiBWarp0) >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t st = it % kStages, sph = (it / kStages) & 1;
      if (ch == 0 && quad == 0 && lane == 0) trace_ev(p, it, 11);
      if (ch == 0)
        epi_b<0>(c, trow, bar, it & 1, st, sph, lane, quad, xs, a1s + st * kA1Bytes);
      else
        epi_b<1>(c, trow, bar, it & 1, st, sph, lane, quad, xs, a1s + st * kA1Bytes);
      if (ch == 0 && quad == 0 && lane == 0) trace_ev(p, it, 9);
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ====================================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar + kBarW, 2 * L::w_bytes);
      for (int off = 0; off < L::w_bytes; off += 16384) {
        bulk_g2s(w1s + off, reinterpret_cast<const uint8_t *>(p.w1_img) + off, 16384, bar + kBarW);
        bulk_g2s(w2s + off, reinterpret_cast<const uint8_t *>(p.w2_img) + off, 16384, bar + kBarW);
      }
      mbar_wait(bar + kBarW, 0);
      constexpr uint32_t idesc1 = idesc_f16(kTileM, H);
      constexpr uint32_t idesc2 = idesc_f16(kTileM, kHidden);
      const uint32_t w1a = smem_u32(w1s), w2a = smem_u32(w2s);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1, ph = it & 1;
        const uint32_t a1a = smem_u32(a1s) + s * kA1Bytes;
        mbar_wait(bar + kBarA1Full + s, ph2);
        tc_fence_after();
        trace_ev(p, it, 2);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int kb = kk >> 2, k = kk & 3;
            const uint64_t da = smem_desc_sw128(a1a + kb * kTileBytes + k * 32);
            const uint64_t db = smem_desc_sw128(w1a + kb * (HID * 128) + half * (H * 128) + k * 32);
            mma_f16_ss(tmem + half * H, da, db, idesc1, kk != 0);
          }
          mma_commit(bar + (half ? kBarD1bFull : kBarD1aFull));
        }
        mma_commit(bar + kBarA1Empty + s);
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
          const uint64_t db = smem_desc_sw128(w2a + (kk >> 2) * kTileBytes + (kk & 3) * 32);
          mma_f16_ts(tmem + kD2Col, tmem + kA2Col + kk * 8, db, idesc2, kk != 0);
        }
        mma_commit(bar + kBarD2Full);
        trace_ev(p, it, 8);
      }
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    // ============================ z tile loader (TMA) ============================
    if (lane == 0) {
      prefetch_tmap(&maps.z);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        uint8_t *a1 = a1s + s * kA1Bytes;
        mbar_wait(bar + kBarStageFree + s, ph2 ^ 1);      // the previous tenant's output has left
        trace_ev(p, it, 0);
        mbar_arrive_expect_tx(bar + kBarA1Full + s, kA1Bytes);
        const int row0 = int(tile * kTileM);
        tma_load_2d(a1, &maps.z, 0, row0, bar + kBarA1Full + s);
        tma_load_2d(a1 + kTileBytes, &maps.z, 64, row0, bar + kBarA1Full + s);
      }
    }
    __syncwarp();
  } else if (warp == kResWarp) {
    // ============ residual rows into the stage, once GEMM 1 has consumed its z (TMA) ===========
    if (lane == 0) {
      prefetch_tmap(&maps.res);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        uint8_t *a1 = a1s + s * kA1Bytes;
        const int row0 = int(tile * kTileM);
        mbar_wait(bar + kBarA1Empty + s, ph2);
        trace_ev(p, it, 1);
        for (int j = 0; j < 2; ++j) {
          mbar_arrive_expect_tx(bar + kBarRFull + 2 * s + j, kTileBytes);
          tma_load_2d(a1 + j * kTileBytes, &maps.res, j * 64, row0, bar + kBarRFull + 2 * s + j);
        }
      }
    }
    __syncwarp();
  } else {
    // ============================ output store (TMA) ==============================
    if (lane == 0) {
      prefetch_tmap(&maps.out);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        const uint8_t *a1 = a1s + s * kA1Bytes;
        const int row0 = int(tile * kTileM);
        for (int j = 0; j < 2; ++j) {
          mbar_wait(bar + kBarOReady + 2 * s + j, ph2);
          tma_store_2d(&maps.out, j * 64, row0, a1 + j * kTileBytes);
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
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace v6k

long long *g_k2_trace = nullptr;   // shared with gfx_umma7.cu

int umma6_mlp_ln_residual(const gfx_model *m, int layer, const __half *z, const __half *h,
                          int64_t n, __half *h_out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) |
       reinterpret_cast<uintptr_t>(h_out)) & 15)
    return fail(GFX_ERR_ARGUMENT, "tcgen05 MLP: activation buffers must be 16-byte aligned");
  v6k::Maps maps;
  int rc = tma::make_rows128_map(&maps.z, z, n, v6k::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.res, h, n, v6k::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.out, h_out, n, v6k::kTileM);
  if (rc) return rc;
  v6k::Consts c;
  const gfx_host_vectors &hv = m->host;
  for (int i = 0; i < kMlpHidden; ++i) c.b1[i] = hv.b1[size_t(layer) * kMlpHidden + i];
  for (int i = 0; i < kHidden; ++i) {
    c.b2[i] = hv.b2[size_t(layer) * kHidden + i];
    c.g[i] = hv.ln_g[size_t(layer) * kHidden + i];
    c.b[i] = hv.ln_b[size_t(layer) * kHidden + i];
  }
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  v6k::Args a{};
  a.w1_img = m->w1_img + wi; a.w2_img = m->w2_img + wi;
  a.n = n;
  a.trace = g_k2_trace;
  GFX_CUDA(cudaFuncSetAttribute(v6k::umma6_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                v6k::Smem::total));
  const int64_t tiles = (n + v6k::kTileM - 1) / v6k::kTileM;
  const int grid = int(tiles < kNumSMs ? tiles : kNumSMs);
  v6k::umma6_mlp_kernel<<<grid, v6k::kWarps * 32, v6k::Smem::total, st>>>(maps, c, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx

// developer hook (not part of include/gfx.h), see gfx_debug_fused_trace
extern "C" int gfx_debug_k2_trace(long long *device_buffer) {
  gfx::g_k2_trace = device_buffer;
  return GFX_OK;
}
