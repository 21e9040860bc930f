// K2 (MLP + LayerNorm + residual), fourth tcgen05 version: the epilogues are
// what the tensor pipe waits for, so they are made lean and wide.
//
// Profile of gfx_umma3.cu (profiles/r01_b): tensor pipe 29 % active, 31 % of
// warp stalls are instruction-cache misses (59 KB of SASS, five roles), the
// two 4-warp epilogues execute ~8,600 warp instructions per 128-row tile with
// two warps per scheduler.  Changes here, same tensor schedule and TMA I/O:
//   * 16 epilogue warps (8 for D1 -> A2, 8 for D2 -> LayerNorm): two warps
//     share a TMEM lane quadrant and split its columns, four warps per
//     scheduler hide each other's latencies;
//   * bias / LayerNorm vectors travel as kernel parameters, so every use is a
//     constant-bank operand of the FADD/FFMA itself (no shared-memory loads);
//   * bias + ReLU + fp16 pack is FADD + cvt.rn.relu.f16x2.f32 (1.5
//     instructions per hidden unit instead of 3.5);
//   * LayerNorm is one pass (sum, sum of squares) with the two column halves'
//     partial sums exchanged through shared memory, then two FFMAs per output;
//   * ~21 KB of hot code.
//
// Warps: 0-7 epilogue A, 8-15 epilogue B, 16 MMA issuer, 17 z loader (TMA),
// 18 residual / output manager (TMA).
//
// Tiles are taken in ascending row order; K1 walks the chunk in DESCENDING order
// (gfx_simt.cu), so what either kernel touched last is what the other one reads
// first while the 126 MB L2 still holds it.
#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v4 {

constexpr int HID = kMlpHidden, H = HID / 2;
constexpr int kTileM = 128;
constexpr int kTileBytes = kTileM * 128;      // [128 x 64] fp16 box
constexpr int kA1Bytes = 2 * kTileBytes;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kA2Col = 256, kD2Col = 384;
constexpr int kEpiBWarp0 = 8, kMmaWarp = 16, kLoadWarp = 17, kIoWarp = 18, kWarps = 19;

enum Bar {
  kBarW = 0, kBarA1Full = 1, kBarA1Empty = 3, kBarD1aFull = 5, kBarD1bFull = 6,
  kBarA2aFull = 7, kBarA2bFull = 8, kBarD2Full = 9, kBarD2Empty = 10,
  kBarRFull = 11, kBarOReady = 13, kNumBars = 15
};

struct Smem {
  static constexpr int w_bytes = HID * kHidden * 2;                 // each weight image
  static constexpr int off_w1 = 0;
  static constexpr int off_w2 = off_w1 + w_bytes;
  static constexpr int off_a1 = off_w2 + w_bytes;                   // 2 stages x 32 KB
  static constexpr int off_r = off_a1 + 2 * kA1Bytes;               // 2 halves x 16 KB
  static constexpr int off_xs = off_r + 2 * kTileBytes;             // float2[2][128] partial sums
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
};

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
                                      int lane, int quad, float2 *xs, uint8_t *rs) {
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
  mbar_wait(bar + kBarRFull + CH, ph);
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
  if (lane == 0) mbar_arrive(bar + kBarOReady + CH);
}

__global__ void __launch_bounds__(kWarps * 32, 1)
umma4_mlp_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Consts c, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2, *a1s = smem + L::off_a1;
  uint8_t *rs = smem + L::off_r;
  float2 *xs = reinterpret_cast<float2 *>(smem + L::off_xs);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::off_bar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
  } else if (tid == 0) {
    mbar_init(bar + kBarW, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar + kBarA1Full + s, 1);
      mbar_init(bar + kBarA1Empty + s, 1);
      mbar_init(bar + kBarRFull + s, 1);
      mbar_init(bar + kBarOReady + s, 4);
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
        epi_a<1, 0>(c, trow, bar, ph, lane);
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
      if (ch == 0)
        epi_b<0>(c, trow, bar, it & 1, lane, quad, xs, rs);
      else
        epi_b<1>(c, trow, bar, it & 1, lane, quad, xs, rs);
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
        const uint32_t s = it & 1, ph2 = (it >> 1) & 1, ph = it & 1;
        const uint32_t a1a = smem_u32(a1s) + s * kA1Bytes;
        mbar_wait(bar + kBarA1Full + s, ph2);
        tc_fence_after();
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
        mbar_wait(bar + kBarA2aFull, ph);
        mbar_wait(bar + kBarD2Empty, ph ^ 1);
        tc_fence_after();
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
      }
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    // ============================ z tile loader (TMA) ============================
    if (lane == 0) {
      prefetch_tmap(&maps.z);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it & 1, ph2 = (it >> 1) & 1;
        uint8_t *a1 = a1s + s * kA1Bytes;
        mbar_wait(bar + kBarA1Empty + s, ph2 ^ 1);
        mbar_arrive_expect_tx(bar + kBarA1Full + s, kA1Bytes);
        const int row0 = int(tile * kTileM);
        tma_load_2d(a1, &maps.z, 0, row0, bar + kBarA1Full + s);
        tma_load_2d(a1 + kTileBytes, &maps.z, 64, row0, bar + kBarA1Full + s);
      }
    }
    __syncwarp();
  } else {
    // ================== residual load / output store manager (TMA) ================
    if (lane == 0) {
      prefetch_tmap(&maps.res);
      prefetch_tmap(&maps.out);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int row0 = int(tile * kTileM);
        if (it == 0) {
          for (int j = 0; j < 2; ++j) {
            mbar_arrive_expect_tx(bar + kBarRFull + j, kTileBytes);
            tma_load_2d(rs + j * kTileBytes, &maps.res, j * 64, row0, bar + kBarRFull + j);
          }
        }
        for (int j = 0; j < 2; ++j) {
          mbar_wait(bar + kBarOReady + j, it & 1);
          tma_store_2d(&maps.out, j * 64, row0, rs + j * kTileBytes);
          bulk_commit();
        }
        const int64_t next = tile + gridDim.x;
        if (next < tiles) {
          bulk_wait_read<1>();          // half 0 has left shared memory
          mbar_arrive_expect_tx(bar + kBarRFull + 0, kTileBytes);
          tma_load_2d(rs, &maps.res, 0, int(next * kTileM), bar + kBarRFull + 0);
          bulk_wait_read<0>();
          mbar_arrive_expect_tx(bar + kBarRFull + 1, kTileBytes);
          tma_load_2d(rs + kTileBytes, &maps.res, 64, int(next * kTileM), bar + kBarRFull + 1);
        }
      }
      bulk_wait_all();                  // every store has landed before the CTA exits
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace v4

int umma4_mlp_ln_residual(const gfx_model *m, int layer, const __half *z, const __half *h,
                          int64_t n, __half *h_out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) |
       reinterpret_cast<uintptr_t>(h_out)) & 15)
    return fail(GFX_ERR_ARGUMENT, "tcgen05 MLP: activation buffers must be 16-byte aligned");
  v4::Maps maps;
  int rc = tma::make_rows128_map(&maps.z, z, n, v4::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.res, h, n, v4::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.out, h_out, n, v4::kTileM);
  if (rc) return rc;
  v4::Consts c;
  const gfx_host_vectors &hv = m->host;
  for (int i = 0; i < kMlpHidden; ++i) c.b1[i] = hv.b1[size_t(layer) * kMlpHidden + i];
  for (int i = 0; i < kHidden; ++i) {
    c.b2[i] = hv.b2[size_t(layer) * kHidden + i];
    c.g[i] = hv.ln_g[size_t(layer) * kHidden + i];
    c.b[i] = hv.ln_b[size_t(layer) * kHidden + i];
  }
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  v4::Args a{};
  a.w1_img = m->w1_img + wi; a.w2_img = m->w2_img + wi;
  a.n = n;
  GFX_CUDA(cudaFuncSetAttribute(v4::umma4_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                v4::Smem::total));
  const int64_t tiles = (n + v4::kTileM - 1) / v4::kTileM;
  const int grid = int(tiles < kNumSMs ? tiles : kNumSMs);
  v4::umma4_mlp_kernel<<<grid, v4::kWarps * 32, v4::Smem::total, st>>>(maps, c, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx
