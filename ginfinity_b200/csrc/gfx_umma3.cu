// K2 (MLP + LayerNorm + residual) with all tile I/O on the TMA engine.
//
// Same tensor-pipe schedule as gfx_umma2.cu (MMA1a | MMA1b | MMA2a | MMA2b with
// the hidden activation kept in TMEM as the A operand of GEMM-2), but no thread
// touches global memory any more:
//   * z tiles (A operand of GEMM-1) arrive by 2-D tiled TMA with the 128-byte
//     swizzle, i.e. already in the UMMA K-major layout (2-stage ring);
//   * the residual tile h arrives the same way into R (two 64-column halves);
//     the final epilogue reads its row chunks from R at the swizzled position
//     (bank-conflict free), adds LayerNorm(u) and writes the result back IN
//     PLACE; a TMA store sends each half to h_out.  Rows past the end of the
//     tensor are zero-filled on load and clipped on store by the tensor map.
// Thread-per-row global accesses cost 32 L1 wavefronts per instruction; this
// removes them (ncu profile r01_a: 2 x 2048 wavefront-cycles per tile).
//
// Warps: 0-3 epilogue A, 4-7 epilogue B, 8 MMA issuer, 9 z loader (1 thread),
// 10 residual/output manager (1 thread).
#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v3 {

constexpr int HID = kMlpHidden, H = HID / 2;
constexpr int kTileM = 128;
constexpr int kTileBytes = kTileM * 128;      // [128 x 64] fp16 box
constexpr int kA1Bytes = 2 * kTileBytes;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kA2Col = 256, kD2Col = 384;
constexpr int kEpiBWarp0 = 4, kMmaWarp = 8, kLoadWarp = 9, kIoWarp = 10, kWarps = 11;

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
  static constexpr int off_b1 = off_r + 2 * kTileBytes;             // float[HID]
  static constexpr int off_vec = off_b1 + HID * 4;                  // float[3][128]
  static constexpr int off_bar = off_vec + 3 * kHidden * 4;
  static constexpr int off_tmem = off_bar + kNumBars * 8;
  static constexpr int total = off_tmem + 8;
};
static_assert(Smem::total <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct alignas(64) Maps {
  CUtensorMap z, res, out;   // [n, 128] fp16, box 64 x 128, SWIZZLE_128B
};

struct Args {
  const __half *w1_img, *w2_img;
  const float *b1, *b2, *ln_g, *ln_b;
  int64_t n;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint4 pack8(const float *v) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4 &raw, float *f) {
  const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(kWarps * 32, 1)
umma3_mlp_kernel(const __grid_constant__ Maps maps, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2, *a1s = smem + L::off_a1;
  uint8_t *rs = smem + L::off_r;
  float *b1s = reinterpret_cast<float *>(smem + L::off_b1);
  float *b2s = reinterpret_cast<float *>(smem + L::off_vec);
  float *gs = b2s + kHidden, *bs = gs + kHidden;
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
    mbar_init(bar + kBarA2aFull, 4);
    mbar_init(bar + kBarA2bFull, 4);
    mbar_init(bar + kBarD2Full, 1);
    mbar_init(bar + kBarD2Empty, 4);
    fence_mbar_init();
  }
  for (int i = tid; i < HID; i += blockDim.x) b1s[i] = p.b1[i];
  for (int i = tid; i < kHidden; i += blockDim.x) {
    b2s[i] = p.b2[i];
    gs[i] = p.ln_g[i];
    bs[i] = p.ln_b[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t tiles = (p.n + kTileM - 1) / kTileM;

  if (warp == kMmaWarp) {
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
  } else if (warp < kEpiBWarp0) {
    // ================= epilogue A: D1 -> bias + ReLU -> fp16 -> A2 (TMEM) =========
    const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(bar + (half ? kBarD1bFull : kBarD1aFull), ph);
        tc_fence_after();
        // all four 32-column loads are issued before the single wait: one
        // TMEM round trip per half instead of four
        float v[H];
#pragma unroll
        for (int cb = 0; cb < H / 32; ++cb) tmem_ld32(trow + half * H + cb * 32, v + cb * 32);
        tmem_ld_wait();
#pragma unroll
        for (int cb = 0; cb < H / 32; ++cb) {
          const int col = half * H + cb * 32;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pk[j] = pack2(fmaxf(v[cb * 32 + 2 * j] + b1s[col + 2 * j], 0.f),
                          fmaxf(v[cb * 32 + 2 * j + 1] + b1s[col + 2 * j + 1], 0.f));
          tmem_st16(trow + kA2Col + col / 2, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + (half ? kBarA2bFull : kBarA2aFull));
      }
    }
  } else if (warp < kMmaWarp) {
    // ===== epilogue B: D2 -> bias + LayerNorm, + residual from R, in place in R =====
    const int wq = warp - kEpiBWarp0;
    const int r = wq * 32 + lane;                          // row inside the tile
    const uint32_t trow = tmem + (uint32_t(wq * 32) << 16) + kD2Col;
    uint8_t *rrow = rs + r * 128;
    const int rx = r & 7;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(bar + kBarD2Full, ph);
      tc_fence_after();
      // copy the whole accumulator row to registers with one TMEM round trip
      // and hand D2 straight back to the MMA warp
      float u[kHidden];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) tmem_ld32(trow + cb * 32, u + cb * 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + kBarD2Empty);
      const float shift = u[0] + b2s[0];
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kHidden; ++j) {
        u[j] = u[j] + b2s[j] - shift;
        s1[j & 3] += u[j];
        s2[j & 3] = fmaf(u[j], u[j], s2[j & 3]);
      }
      const float m = ((s1[0] + s1[1]) + (s1[2] + s1[3])) * (1.f / kHidden);
      const float var = fmaxf(((s2[0] + s2[1]) + (s2[2] + s2[3])) * (1.f / kHidden) - m * m, 0.f);
      const float mul = rsqrtf(var + 1e-5f);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        if ((cb & 1) == 0) mbar_wait(bar + kBarRFull + (cb >> 1), ph);
        uint8_t *half_base = rrow + (cb >> 1) * kTileBytes;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 *cell = reinterpret_cast<uint4 *>(half_base + ((((cb & 1) * 4 + g) ^ rx) << 4));
          float r8[8], o[8];
          unpack8(*cell, r8);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = cb * 32 + g * 8 + j;
            o[j] = r8[j] + ((u[c] - m) * mul * gs[c] + bs[c]);
          }
          *cell = pack8(o);
        }
        if (cb & 1) {                   // this 64-column half of the tile is final
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar + kBarOReady + (cb >> 1));
        }
      }
    }
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
        tma_load_2d(a1, &maps.z, 0, int(tile * kTileM), bar + kBarA1Full + s);
        tma_load_2d(a1 + kTileBytes, &maps.z, 64, int(tile * kTileM), bar + kBarA1Full + s);
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

}  // namespace v3

int umma3_mlp_ln_residual(const gfx_model *m, int layer, const __half *z, const __half *h,
                          int64_t n, __half *h_out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) |
       reinterpret_cast<uintptr_t>(h_out)) & 15)
    return fail(GFX_ERR_ARGUMENT, "tcgen05 MLP: activation buffers must be 16-byte aligned");
  v3::Maps maps;
  int rc = tma::make_rows128_map(&maps.z, z, n, v3::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.res, h, n, v3::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.out, h_out, n, v3::kTileM);
  if (rc) return rc;
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  v3::Args a{};
  a.w1_img = m->w1_img + wi; a.w2_img = m->w2_img + wi;
  a.b1 = m->b1 + size_t(layer) * kMlpHidden; a.b2 = m->b2 + size_t(layer) * kHidden;
  a.ln_g = m->ln_g + size_t(layer) * kHidden; a.ln_b = m->ln_b + size_t(layer) * kHidden;
  a.n = n;
  GFX_CUDA(cudaFuncSetAttribute(v3::umma3_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                v3::Smem::total));
  const int64_t tiles = (n + v3::kTileM - 1) / v3::kTileM;
  const int grid = int(tiles < kNumSMs ? tiles : kNumSMs);
  v3::umma3_mlp_kernel<<<grid, v3::kWarps * 32, v3::Smem::total, st>>>(maps, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx
