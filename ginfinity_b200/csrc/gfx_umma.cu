// K2 / K3 on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM).  fp16 operands, fp32 accumulation.
//
//   K2  h_out = h + LayerNorm(W2 relu(W1' z + b1') + b2)      HID = 256
//   K3  out   = l2norm(Wb relu(Wa h + ba) + bb)               HID = 128
//
// Persistent CTAs (one per SM).  Both weight matrices of the stage stay
// resident in shared memory for the whole kernel as pre-swizzled K-major
// images (built once on the host by gfx_model_create, fetched with bulk
// async copies on the TMA engine).  Per 128-node tile:
//   A1 (128x128 fp16 activations) is staged into the 128B-swizzled K-major
//   layout, GEMM-1 accumulates D1[128 x HID] in TMEM, the epilogue applies
//   bias+ReLU and writes fp16 straight back to shared memory as the A operand
//   of GEMM-2 (the hidden activation never leaves the SM), GEMM-2 accumulates
//   D2[128 x 128] in TMEM, and the final epilogue does bias + LayerNorm +
//   residual (or bias + L2 normalise) with one thread per node row, so the
//   row reductions are thread-local.
#include "gfx_common.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

constexpr int kTileM = 128;               // node rows per tile = UMMA M
constexpr int kKBlock = 64;               // fp16 columns per 128-byte swizzle row
constexpr int kTileBytes = kTileM * 128;  // one [128 x 64] fp16 K-block tile
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kD2Col = 256;          // D1 at columns [0, HID), D2 at [256, 384)

template <int HID>
struct UmmaSmem {
  static constexpr int w1_bytes = HID * kHidden * 2;       // B of GEMM-1: [HID x 128]
  static constexpr int w2_bytes = kHidden * HID * 2;       // B of GEMM-2: [128 x HID]
  static constexpr int a1_bytes = kTileM * kHidden * 2;    // A of GEMM-1
  static constexpr int a2_bytes = kTileM * HID * 2;        // A of GEMM-2
  static constexpr int off_w1 = 0;
  static constexpr int off_w2 = off_w1 + w1_bytes;
  static constexpr int off_a1 = off_w2 + w2_bytes;
  static constexpr int off_a2 = off_a1 + a1_bytes;
  static constexpr int off_b1 = off_a2 + a2_bytes;               // float[HID]
  static constexpr int off_vec = off_b1 + HID * 4;               // float[3][128]: b2, g, b
  static constexpr int off_bar = off_vec + 3 * kHidden * 4;      // 4 x uint64
  static constexpr int off_tmem = off_bar + 4 * 8;               // uint32
  static constexpr int total = off_tmem + 16;
};

__device__ __forceinline__ uint4 pack8(const float *v) {
  uint4 r;
  __half2 *h = reinterpret_cast<__half2 *>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return r;
}

// byte offset of 16-byte chunk `c16` (0..HID/8) of row `r` inside a K-major
// swizzled operand made of [128 x 64] tiles
__device__ __forceinline__ uint32_t a_chunk_offset(int r, int c16) {
  return uint32_t(c16 >> 3) * kTileBytes + uint32_t(r) * 128 + uint32_t(((c16 & 7) ^ (r & 7)) << 4);
}

template <int HID, int MODE, typename TOut>
__global__ void __launch_bounds__(128, 1)
umma_mlp_kernel(const __half *__restrict__ a_in, const __half *__restrict__ res,
                const __half *__restrict__ w1_img, const float *__restrict__ b1,
                const __half *__restrict__ w2_img, const float *__restrict__ b2,
                const float *__restrict__ ln_g, const float *__restrict__ ln_b,
                const int32_t *__restrict__ out_row, int64_t n, TOut *__restrict__ out) {
  using L = UmmaSmem<HID>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2;
  uint8_t *a1s = smem + L::off_a1, *a2s = smem + L::off_a2;
  float *b1s = reinterpret_cast<float *>(smem + L::off_b1);
  float *b2s = reinterpret_cast<float *>(smem + L::off_vec);
  float *gs = b2s + kHidden, *bs = gs + kHidden;
  uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem + L::off_bar);
  uint64_t *bar_mma1 = bar_w + 1, *bar_mma2 = bar_w + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    tmem_alloc(tmem_slot, kTmemCols);
  } else if (tid == 32) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma1, 1);
    mbar_init(bar_mma2, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < HID; i += 128) b1s[i] = b1[i];
  b2s[tid] = b2[tid];
  if (MODE == 0) {
    gs[tid] = ln_g[tid];
    bs[tid] = ln_b[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + (uint32_t(warp * 32) << 16);

  if (tid == 0) {  // weights: two images, fetched once per CTA by the TMA engine
    mbar_arrive_expect_tx(bar_w, L::w1_bytes + L::w2_bytes);
    for (int off = 0; off < L::w1_bytes; off += 16384)
      bulk_g2s(w1s + off, reinterpret_cast<const uint8_t *>(w1_img) + off, 16384, bar_w);
    for (int off = 0; off < L::w2_bytes; off += 16384)
      bulk_g2s(w2s + off, reinterpret_cast<const uint8_t *>(w2_img) + off, 16384, bar_w);
  }

  const int64_t tiles = (n + kTileM - 1) / kTileM;
  uint32_t phase = 0;
  bool first = true;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, phase ^= 1) {
    const int64_t row0 = tile * kTileM;
    // ---- stage A1: 128 rows x 256 bytes, coalesced 16-byte cp.async ------------
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int q = i * 128 + tid, r = q >> 4, c16 = q & 15;
      const bool ok = row0 + r < n;
      const __half *src = a_in + (ok ? (row0 + r) : 0) * kHidden + c16 * 8;
      cp_async16(a1s + a_chunk_offset(r, c16), src, ok ? 16u : 0u);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_async_smem();
    __syncthreads();
    // ---- GEMM-1: D1[128 x HID] = A1[128 x 128] * W1'^T --------------------------
    if (tid == 0) {
      if (first) mbar_wait(bar_w, 0);
      tc_fence_after();
      constexpr uint32_t idesc = idesc_f16(kTileM, HID);
#pragma unroll
      for (int kb = 0; kb < kHidden / kKBlock; ++kb)
#pragma unroll
        for (int k = 0; k < kKBlock / 16; ++k) {
          const uint64_t da = smem_desc_sw128(smem_u32(a1s) + kb * kTileBytes + k * 32);
          const uint64_t db = smem_desc_sw128(smem_u32(w1s) + kb * (HID * 128) + k * 32);
          mma_f16_ss(tmem, da, db, idesc, (kb | k) != 0);
        }
      mma_commit(bar_mma1);
    }
    first = false;
    // ---- epilogue 1: bias + ReLU -> fp16 -> A2 (swizzled K-major) ---------------
    mbar_wait(bar_mma1, phase);
    tc_fence_after();
#pragma unroll
    for (int cb = 0; cb < HID / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_row + cb * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + b1s[cb * 32 + j], 0.f);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4 *>(a2s + a_chunk_offset(tid, cb * 4 + g)) = pack8(v + g * 8);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- GEMM-2: D2[128 x 128] = A2[128 x HID] * W2^T ---------------------------
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = idesc_f16(kTileM, kHidden);
#pragma unroll
      for (int kb = 0; kb < HID / kKBlock; ++kb)
#pragma unroll
        for (int k = 0; k < kKBlock / 16; ++k) {
          const uint64_t da = smem_desc_sw128(smem_u32(a2s) + kb * kTileBytes + k * 32);
          const uint64_t db = smem_desc_sw128(smem_u32(w2s) + kb * kTileBytes + k * 32);
          mma_f16_ss(tmem + kD2Col, da, db, idesc, (kb | k) != 0);
        }
      mma_commit(bar_mma2);
    }
    // ---- epilogue 2: one thread per node row ------------------------------------
    mbar_wait(bar_mma2, phase);
    tc_fence_after();
    float u[kHidden];
#pragma unroll
    for (int cb = 0; cb < kHidden / 32; ++cb) tmem_ld32(tmem_row + kD2Col + cb * 32, u + cb * 32);
    tmem_ld_wait();
    tc_fence_before();
    const int64_t row = row0 + tid;
    if (MODE == 0) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < kHidden; ++j) {
        u[j] += b2s[j];
        s += u[j];
      }
      const float mean = s * (1.f / kHidden);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < kHidden; ++j) {
        const float d = u[j] - mean;
        q = fmaf(d, d, q);
      }
      const float rstd = rsqrtf(q * (1.f / kHidden) + 1e-5f);
      if (row < n) {
        const uint4 *rp = reinterpret_cast<const uint4 *>(res + row * kHidden);
        uint4 *op = reinterpret_cast<uint4 *>(out + row * kHidden);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint4 raw = rp[c];
          const __half2 *hh = reinterpret_cast<const __half2 *>(&raw);
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = c * 8 + j;
            const float2 f = __half22float2(hh[j >> 1]);
            const float hres = (j & 1) ? f.y : f.x;
            o[j] = hres + ((u[col] - mean) * rstd * gs[col] + bs[col]);
          }
          op[c] = pack8(o);
        }
      }
    } else {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < kHidden; ++j) {
        u[j] += b2s[j];
        q = fmaf(u[j], u[j], q);
      }
      const float inv = 1.f / fmaxf(sqrtf(q), 1e-12f);
      if (row < n) {
        const int64_t orow = out_row ? int64_t(out_row[row]) : row;
        if (orow >= 0) {
          if (sizeof(TOut) == 2) {
            uint4 *op = reinterpret_cast<uint4 *>(out + orow * kHidden);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = u[c * 8 + j] * inv;
              op[c] = pack8(o);
            }
          } else {
            float4 *op = reinterpret_cast<float4 *>(out + orow * kHidden);
#pragma unroll
            for (int c = 0; c < 32; ++c)
              op[c] = make_float4(u[4 * c] * inv, u[4 * c + 1] * inv, u[4 * c + 2] * inv,
                                  u[4 * c + 3] * inv);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

template <int HID, int MODE, typename TOut>
static int launch_umma(const __half *a, const __half *res, const __half *w1, const float *b1,
                       const __half *w2, const float *b2, const float *g, const float *b,
                       const int32_t *out_row, int64_t n, TOut *out, cudaStream_t st) {
  auto kern = umma_mlp_kernel<HID, MODE, TOut>;
  constexpr int smem = UmmaSmem<HID>::total;
  static_assert(smem <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
  GFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (n + kTileM - 1) / kTileM;
  const int grid = int(tiles < kNumSMs ? tiles : kNumSMs);
  kern<<<grid, 128, smem, st>>>(a, res, w1, b1, w2, b2, g, b, out_row, n, out);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

int umma1_mlp_ln_residual(const gfx_model *m, int layer, const __half *z, const __half *h,
                         int64_t n, __half *h_out, cudaStream_t st) {
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  return launch_umma<kMlpHidden, 0, __half>(
      z, h, m->w1_img + wi, m->b1 + size_t(layer) * kMlpHidden, m->w2_img + wi,
      m->b2 + size_t(layer) * kHidden, m->ln_g + size_t(layer) * kHidden,
      m->ln_b + size_t(layer) * kHidden, nullptr, n, h_out, st);
}

int umma1_head_l2norm(const gfx_model *m, const __half *h, const int32_t *out_row, int64_t n,
                     void *out, int out_dtype, cudaStream_t st) {
  if (out_dtype == GFX_F16)
    return launch_umma<kHidden, 1, __half>(h, nullptr, m->wa_img, m->ba, m->wb_img, m->bb, nullptr,
                                           nullptr, out_row, n, static_cast<__half *>(out), st);
  return launch_umma<kHidden, 1, float>(h, nullptr, m->wa_img, m->ba, m->wb_img, m->bb, nullptr,
                                        nullptr, out_row, n, static_cast<float *>(out), st);
}

}  // namespace gfx

