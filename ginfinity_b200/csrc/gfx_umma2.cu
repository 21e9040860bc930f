// K2 / K3 on tcgen05: warp-specialised, software-pipelined (second version; K2's
// current kernel is gfx_umma4.cu, the fused layer is gfx_fused5.cu).
//
//   K2     h_out = h + LayerNorm(W2 relu(W1' z + b1') + b2)          HID = 256
//   K3     out   = l2norm(Wb relu(Wa h + ba) + bb)                   HID = 128
//
// One persistent CTA per SM.  Roles (warps):
//   0-3   epilogue A   D1 (TMEM) -> bias + ReLU -> fp16 -> A2 (TMEM)
//   4-7   epilogue B   D2 (TMEM) -> bias + LayerNorm + residual | L2 norm -> HBM
//   8     MMA issuer   one thread issues every tcgen05.mma; also owns TMEM
//                      allocation and the one-time weight fetch (TMA bulk copy)
//   9..   producers    fill the A1 ring (2 stages x 128 rows x 128 fp16,
//                      K-major, 128-byte swizzle): cp.async of the input rows
//
// Tensor-pipe schedule per 128-node tile (H = HID/2):
//     MMA1a  D1[:, :H]  = A1 * W1[:H]^T      | epilogue A on the previous half
//     MMA1b  D1[:, H:]  = A1 * W1[H:]^T      | epilogue A on D1[:, :H]
//     MMA2a  D2  = A2[:, :H] * W2[:, :H]^T   | epilogue A on D1[:, H:]
//     MMA2b  D2 += A2[:, H:] * W2[:, H:]^T   | epilogue B on the previous tile
// The hidden activation A2 is written to tensor memory and consumed from
// there (tcgen05.mma with A in TMEM), so it costs neither shared-memory space
// nor shared-memory bandwidth; TMEM columns: D1 [0,HID) A2 [256,256+HID/2)
// D2 [384,512).
#include "gfx_common.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v2 {

constexpr int kTileM = 128;
constexpr int kTileBytes = kTileM * 128;      // one [128 x 64] fp16 K-block tile (16 KB)
constexpr int kA1Bytes = 2 * kTileBytes;      // [128 x 128] fp16
constexpr int kStages = 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kA2Col = 256, kD2Col = 384;
constexpr int kEpiAWarp0 = 0, kEpiBWarp0 = 4, kMmaWarp = 8, kProdWarp0 = 9;

enum Bar {
  kBarW = 0, kBarA1Full = 1, kBarA1Empty = 3, kBarD1aFull = 5, kBarD1bFull = 6,
  kBarA2aFull = 7, kBarA2bFull = 8, kBarD2Full = 9, kBarD2Empty = 10,
  kNumBars = 11
};

template <int HID>
struct Smem {
  static constexpr int w1_bytes = HID * kHidden * 2;
  static constexpr int w2_bytes = kHidden * HID * 2;
  static constexpr int off_w1 = 0;
  static constexpr int off_w2 = off_w1 + w1_bytes;
  static constexpr int off_a1 = off_w2 + w2_bytes;                 // kStages x 32 KB
  static constexpr int off_b1 = off_a1 + kStages * kA1Bytes;       // float[HID]
  static constexpr int off_vec = off_b1 + HID * 4;                 // float[3][128]
  static constexpr int off_bar = off_vec + 3 * kHidden * 4;
  static constexpr int off_tmem = off_bar + kNumBars * 8;
  static constexpr int total = off_tmem + 8;
};

struct Args {
  const __half *a_in;     // the GEMM-1 input rows
  const __half *res;      // residual rows (MODE 0)
  const __half *w1_img, *w2_img;
  const float *b1, *b2, *ln_g, *ln_b;
  const int32_t *out_row;
  int64_t n;
  void *out;
};

__device__ __forceinline__ uint32_t a_chunk_offset(int r, int c16) {
  return uint32_t(c16 >> 3) * kTileBytes + uint32_t(r) * 128 + uint32_t(((c16 & 7) ^ (r & 7)) << 4);
}

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

template <int HID, int MODE, typename TOut, int NPROD>
__global__ void __launch_bounds__((kProdWarp0 + NPROD) * 32, 1)
umma2_kernel(const Args p) {
  using L = Smem<HID>;
  constexpr int H = HID / 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2, *a1s = smem + L::off_a1;
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
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar + kBarA1Full + s, NPROD);
      mbar_init(bar + kBarA1Empty + s, 1);
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
    if (MODE == 0) {
      gs[i] = p.ln_g[i];
      bs[i] = p.ln_b[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t tiles = (p.n + kTileM - 1) / kTileM;

  if (warp == kMmaWarp) {
    // ======================= MMA issuer (one thread) ==========================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar + kBarW, L::w1_bytes + L::w2_bytes);
      for (int off = 0; off < L::w1_bytes; off += 16384)
        bulk_g2s(w1s + off, reinterpret_cast<const uint8_t *>(p.w1_img) + off, 16384, bar + kBarW);
      for (int off = 0; off < L::w2_bytes; off += 16384)
        bulk_g2s(w2s + off, reinterpret_cast<const uint8_t *>(p.w2_img) + off, 16384, bar + kBarW);
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
        mma_commit(bar + kBarA1Empty + s);     // both halves have consumed this A1 stage
        // GEMM-2, first K half: needs A2[:, :H] (epilogue A) and a drained D2
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
    // ================= epilogue A: D1 -> relu -> fp16 -> A2 (TMEM) ===============
    const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(bar + (half ? kBarD1bFull : kBarD1aFull), ph);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < H / 32; ++cb) {
          const int col = half * H + cb * 32;
          float v[32];
          tmem_ld32(trow + col, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pk[j] = pack2(fmaxf(v[2 * j] + b1s[col + 2 * j], 0.f),
                          fmaxf(v[2 * j + 1] + b1s[col + 2 * j + 1], 0.f));
          tmem_st16(trow + kA2Col + col / 2, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + (half ? kBarA2bFull : kBarA2aFull));
      }
    }
  } else if (warp < kMmaWarp) {
    // ============ epilogue B: D2 -> LayerNorm + residual | L2 norm -> HBM ==========
    const int wq = warp - kEpiBWarp0;
    const uint32_t trow = tmem + (uint32_t(wq * 32) << 16) + kD2Col;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int64_t row = tile * kTileM + wq * 32 + lane;
      mbar_wait(bar + kBarD2Full, it & 1);
      tc_fence_after();
      // pass 1: statistics.  Shifting by the first element keeps the one-pass
      // variance free of cancellation.
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld32(trow + cb * 32, v);
        tmem_ld_wait();
        if (cb == 0) shift = MODE == 0 ? v[0] + b2s[0] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = v[j] + b2s[cb * 32 + j] - shift;
          s1 += d;
          s2 = fmaf(d, d, s2);
        }
      }
      float mul, sub;   // out = (x - sub) * mul [* g + b]
      if (MODE == 0) {
        const float m = s1 * (1.f / kHidden);
        const float var = fmaxf(s2 * (1.f / kHidden) - m * m, 0.f);
        mul = rsqrtf(var + 1e-5f);
        sub = shift + m;
      } else {
        mul = 1.f / fmaxf(sqrtf(s2), 1e-12f);
        sub = 0.f;
      }
      const bool live = row < p.n;
      int64_t orow = row;
      if (MODE == 1 && live && p.out_row) orow = p.out_row[row];
      const bool store = live && orow >= 0;
      // pass 2: normalise and store
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld32(trow + cb * 32, v);
        tmem_ld_wait();
        if (cb == 3) {              // D2 fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar + kBarD2Empty);
        }
        if (!store) continue;
        if (MODE == 0) {
          const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + row * kHidden + cb * 32);
          uint4 *op = reinterpret_cast<uint4 *>(static_cast<__half *>(p.out) + row * kHidden + cb * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float r8[8], o[8];
            unpack8(rp[g], r8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = cb * 32 + g * 8 + j;
              o[j] = r8[j] + ((v[g * 8 + j] + b2s[c] - sub) * mul * gs[c] + bs[c]);
            }
            op[g] = pack8(o);
          }
        } else if (sizeof(TOut) == 2) {
          // a thread holds one row: 256-bit stores (a whole sector per thread) halve the L1
          // wavefronts of these 32-lines-per-warp accesses (gfx_split9.cu)
          __half *op = static_cast<__half *>(p.out) + orow * kHidden + cb * 32;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = (v[g * 16 + j] + b2s[cb * 32 + g * 16 + j]) * mul;
            const uint4 lo = pack8(o), hi = pack8(o + 8);
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op + g * 16),
                         "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z),
                         "r"(hi.w)
                         : "memory");
          }
        } else {
          float *op = static_cast<float *>(p.out) + orow * kHidden + cb * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[g * 8 + j] + b2s[cb * 32 + g * 8 + j]) * mul;
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op + g * 8),
                         "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]),
                         "f"(o[7])
                         : "memory");
          }
        }
      }
    }
  } else {
    // ================================ producers ===================================
    const int pw = warp - kProdWarp0;
    const int ptid = pw * 32 + lane;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t s = it & 1, ph2 = (it >> 1) & 1;
      const int64_t row0 = tile * kTileM;
      uint8_t *a1 = a1s + s * kA1Bytes;
      mbar_wait(bar + kBarA1Empty + s, ph2 ^ 1);
      for (int q = ptid; q < kTileM * 16; q += NPROD * 32) {
        const int r = q >> 4, c16 = q & 15;
        const bool ok = row0 + r < p.n;
        const __half *src = p.a_in + (ok ? (row0 + r) : 0) * kHidden + c16 * 8;
        cp_async16(a1 + a_chunk_offset(r, c16), src, ok ? 16u : 0u);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + kBarA1Full + s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

template <int HID, int MODE, typename TOut, int NPROD>
static int launch(const Args &args, cudaStream_t st) {
  auto kern = umma2_kernel<HID, MODE, TOut, NPROD>;
  constexpr int smem = Smem<HID>::total;
  static_assert(smem <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
  GFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (args.n + kTileM - 1) / kTileM;
  const int grid = int(tiles < kNumSMs ? tiles : kNumSMs);
  kern<<<grid, (kProdWarp0 + NPROD) * 32, smem, st>>>(args);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace v2

int umma_mlp_ln_residual(const gfx_model *m, int layer, const __half *z, const __half *h,
                         int64_t n, __half *h_out, cudaStream_t st) {
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  v2::Args a{};
  a.a_in = z; a.res = h; a.w1_img = m->w1_img + wi; a.w2_img = m->w2_img + wi;
  a.b1 = m->b1 + size_t(layer) * kMlpHidden; a.b2 = m->b2 + size_t(layer) * kHidden;
  a.ln_g = m->ln_g + size_t(layer) * kHidden; a.ln_b = m->ln_b + size_t(layer) * kHidden;
  a.n = n; a.out = h_out;
  return v2::launch<kMlpHidden, 0, __half, 4>(a, st);
}

int umma_head_l2norm(const gfx_model *m, const __half *h, const int32_t *out_row, int64_t n,
                     void *out, int out_dtype, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(out) & 31)
    return fail(GFX_ERR_ARGUMENT, "tcgen05 head: the output table must be 32-byte aligned");
  v2::Args a{};
  a.a_in = h; a.w1_img = m->wa_img; a.w2_img = m->wb_img; a.b1 = m->ba; a.b2 = m->bb;
  a.out_row = out_row; a.n = n; a.out = out;
  if (out_dtype == GFX_F16) return v2::launch<kHidden, 1, __half, 4>(a, st);
  return v2::launch<kHidden, 1, float, 4>(a, st);
}

}  // namespace gfx
