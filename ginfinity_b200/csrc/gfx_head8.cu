// K3 (head 128 -> 128 -> 128 + L2 normalise + fp16 cast), TMA version for the common case: fp16 in,
// fp16 out, every node a core node (no row compaction).
//
//   y = Wb relu(Wa h + ba) + bb ;  out = y / max(|y|, 1e-12)          (_model.py:61-63,72; api.py:250-252)
//
// gfx_umma2.cu's head gathers its input with 16-byte cp.async pieces issued by four producer
// warps and lets every epilogue thread write its own 256-byte row: 2.6-2.9 ms per 2.0e7-node pass,
// 0.42 of the HBM peak for 512 B per node.  Here the tile I/O is two TMA loads and two TMA stores
// per 128-row tile, and a stage buffer cycles h tile -> output tile (the MMA has consumed the input
// by the time epilogue B has the norms), four stages deep, as gfx_umma6.cu does for K2.
//
// TMEM: D1 [0,128) fp32, A2 [128,192) fp16 hidden activation (GEMM 2's A operand), D2 [256,384).
// Warps: 0-7 epilogue A (bias + ReLU -> A2), 8-15 epilogue B (bias, row norm with the two column
// halves' sums exchanged through shared memory, scale, fp16, in place in the stage buffer),
// 16 MMA issuer, 17 loader (TMA), 18 store (TMA).
#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v8h {

constexpr int kTileM = 128;
constexpr int kKbBytes = kTileM * 128;        // one K block of a tile / of a weight image: [128 x 64] fp16
constexpr int kTileBytes = 2 * kKbBytes;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kA2Col = 128, kD2Col = 256;
constexpr int kEpiBWarp0 = 8, kMmaWarp = 16, kLoadWarp = 17, kStoreWarp = 18, kWarps = 19;
constexpr int kStages = 4;

enum Bar {
  kBarW = 0, kBarD1Full = 1, kBarA2Full = 2, kBarD2Full = 3, kBarD2Empty = 4,
  kBarA1Full = 5,                          // [stage]     h tile landed
  kBarA1Empty = kBarA1Full + kStages,      // [stage]     GEMM 1 has consumed it
  kBarStageFree = kBarA1Empty + kStages,   // [stage]     the store has read the output
  kBarOReady = kBarStageFree + kStages,    // [stage][2]  output half written
  kNumBars = kBarOReady + 2 * kStages
};

struct Smem {
  static constexpr int off_wa = 0;                                   // 32 KB
  static constexpr int off_wb = off_wa + kTileBytes;                 // 32 KB
  static constexpr int off_a1 = off_wb + kTileBytes;                 // 4 stages x 32 KB
  static constexpr int off_xs = off_a1 + kStages * kTileBytes;       // float[2][128] partial sums of squares
  static constexpr int off_bar = off_xs + 2 * kTileM * 4;
  static constexpr int off_tmem = off_bar + kNumBars * 8;
  static constexpr int total = off_tmem + 8;
};
static_assert(Smem::total <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct alignas(64) Maps {
  CUtensorMap in, out;       // [n, 128] fp16, box 64 x 128, SWIZZLE_128B
};

struct Consts {              // kernel parameters = constant bank: free ALU operands
  float ba[kHidden], bb[kHidden];
};

struct Args {
  const __half *wa_img, *wb_img;
  int64_t n;
};

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

// D1[:, 64 CH .. +64) -> bias + ReLU -> fp16 -> A2
template <int CH>
__device__ __forceinline__ void epi_a(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph, int lane) {
  constexpr int col0 = CH * 64;
  mbar_wait(bar + kBarD1Full, ph);
  tc_fence_after();
  float v[64];
  tmem_ld32(trow + col0, v);
  tmem_ld32(trow + col0 + 32, v + 32);
  tmem_ld_wait();
  uint32_t pk[32];
#pragma unroll
  for (int j = 0; j < 32; ++j)
    pk[j] = relu_pack2(v[2 * j] + c.ba[col0 + 2 * j], v[2 * j + 1] + c.ba[col0 + 2 * j + 1]);
  tmem_st16(trow + kA2Col + col0 / 2, pk);
  tmem_st16(trow + kA2Col + col0 / 2 + 16, pk + 16);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar + kBarA2Full);
}

// D2[:, 64 CH .. +64) -> + bb -> row norm over all 128 columns -> scale -> fp16 into K block CH of
// the stage buffer (whose h tile GEMM 1 has consumed)
template <int CH>
__device__ __forceinline__ void epi_b(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      uint32_t st, uint32_t sph, int lane, int quad, float *xs,
                                      uint8_t *stage) {
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
  float s2[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    u[j] += c.bb[col0 + j];
    s2[j & 1] = fmaf(u[j], u[j], s2[j & 1]);
  }
  xs[CH * kTileM + r] = s2[0] + s2[1];
  named_bar_sync(1 + quad, 64);
  const float other = xs[(CH ^ 1) * kTileM + r];
  named_bar_sync(1 + quad, 64);                          // xs may be rewritten for the next tile
  const float total = CH == 0 ? (s2[0] + s2[1]) + other : other + (s2[0] + s2[1]);   // same sum in both halves
  const float mul = 1.f / fmaxf(sqrtf(total), 1e-12f);
  mbar_wait(bar + kBarA1Empty + st, sph);                // the MMA has read the h tile of this stage
  uint8_t *rrow = stage + CH * kKbBytes + r * 128;
  const int rx = r & 7;
#pragma unroll
  for (int gi = 0; gi < 8; ++gi) {
    uint4 *cell = reinterpret_cast<uint4 *>(rrow + ((gi ^ rx) << 4));
    const int j = gi * 8;
    *cell = make_uint4(pack2(u[j] * mul, u[j + 1] * mul), pack2(u[j + 2] * mul, u[j + 3] * mul),
                       pack2(u[j + 4] * mul, u[j + 5] * mul), pack2(u[j + 6] * mul, u[j + 7] * mul));
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar + kBarOReady + 2 * st + CH);
}

__global__ void __launch_bounds__(kWarps * 32, 1)
head8_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Consts c, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *was = smem + L::off_wa, *wbs = smem + L::off_wb, *a1s = smem + L::off_a1;
  float *xs = reinterpret_cast<float *>(smem + L::off_xs);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::off_bar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
  } else if (tid == 0) {
    mbar_init(bar + kBarW, 1);
    mbar_init(bar + kBarD1Full, 1);
    mbar_init(bar + kBarA2Full, 8);
    mbar_init(bar + kBarD2Full, 1);
    mbar_init(bar + kBarD2Empty, 8);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar + kBarA1Full + s, 1);
      mbar_init(bar + kBarA1Empty + s, 1);
      mbar_init(bar + kBarStageFree + s, 1);
      mbar_init(bar + kBarOReady + 2 * s, 4);
      mbar_init(bar + kBarOReady + 2 * s + 1, 4);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t tiles = (p.n + kTileM - 1) / kTileM;

  if (warp < kEpiBWarp0) {
    // ================= epilogue A =================================================
    const int quad = warp & 3, ch = warp >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      if (ch == 0)
        epi_a<0>(c, trow, bar, it & 1, lane);
      else
        epi_a<1>(c, trow, bar, it & 1, lane);
    }
  } else if (warp < kMmaWarp) {
    // ================= epilogue B =================================================
    const int quad = warp & 3, ch = (warp - kEpiBWarp0) >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t st = it % kStages, sph = (it / kStages) & 1;
      if (ch == 0)
        epi_b<0>(c, trow, bar, it & 1, st, sph, lane, quad, xs, a1s + st * kTileBytes);
      else
        epi_b<1>(c, trow, bar, it & 1, st, sph, lane, quad, xs, a1s + st * kTileBytes);
    }
  } else if (warp == kMmaWarp) {
    // ================= weights + MMA issue ========================================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar + kBarW, 2 * kTileBytes);
      for (int off = 0; off < kTileBytes; off += 16384) {
        bulk_g2s(was + off, reinterpret_cast<const uint8_t *>(p.wa_img) + off, 16384, bar + kBarW);
        bulk_g2s(wbs + off, reinterpret_cast<const uint8_t *>(p.wb_img) + off, 16384, bar + kBarW);
      }
      mbar_wait(bar + kBarW, 0);
      constexpr uint32_t idesc = idesc_f16(kTileM, kHidden);
      const uint32_t waa = smem_u32(was), wba = smem_u32(wbs);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1, ph = it & 1;
        const uint32_t a1a = smem_u32(a1s) + s * kTileBytes;
        mbar_wait(bar + kBarA1Full + s, ph2);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int kb = kk >> 2, k = kk & 3;
          const uint64_t da = smem_desc_sw128(a1a + kb * kKbBytes + k * 32);
          const uint64_t db = smem_desc_sw128(waa + kb * kKbBytes + k * 32);
          mma_f16_ss(tmem, da, db, idesc, kk != 0);
        }
        mma_commit(bar + kBarD1Full);
        mma_commit(bar + kBarA1Empty + s);
        mbar_wait(bar + kBarA2Full, ph);
        mbar_wait(bar + kBarD2Empty, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t db = smem_desc_sw128(wba + (kk >> 2) * kKbBytes + (kk & 3) * 32);
          mma_f16_ts(tmem + kD2Col, tmem + kA2Col + kk * 8, db, idesc, kk != 0);
        }
        mma_commit(bar + kBarD2Full);
      }
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    // ================= h tile loader (TMA) ========================================
    if (lane == 0) {
      prefetch_tmap(&maps.in);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        uint8_t *a1 = a1s + s * kTileBytes;
        mbar_wait_parked(bar + kBarStageFree + s, ph2 ^ 1);   // the previous tenant's output has left
        mbar_arrive_expect_tx(bar + kBarA1Full + s, kTileBytes);
        const int row0 = int(tile * kTileM);
        tma_load_2d(a1, &maps.in, 0, row0, bar + kBarA1Full + s);
        tma_load_2d(a1 + kKbBytes, &maps.in, 64, row0, bar + kBarA1Full + s);
      }
    }
    __syncwarp();
  } else {
    // ================= output store (TMA) =========================================
    if (lane == 0) {
      prefetch_tmap(&maps.out);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages, ph2 = (it / kStages) & 1;
        const uint8_t *a1 = a1s + s * kTileBytes;
        const int row0 = int(tile * kTileM);
        for (int j = 0; j < 2; ++j) {
          mbar_wait_parked(bar + kBarOReady + 2 * s + j, ph2);
          tma_store_2d(&maps.out, j * 64, row0, a1 + j * kKbBytes);
        }
        bulk_commit();
        bulk_wait_read<0>();            // shared memory has been read: the stage takes its next tile
        mbar_arrive(bar + kBarStageFree + s);
      }
      bulk_wait_all();                  // every store has landed before the CTA exits
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace v8h

// fp16 in, fp16 out, identity row map
int head8_l2norm(const gfx_model *m, const __half *h, int64_t n, __half *out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(GFX_ERR_ARGUMENT, "tcgen05 head: buffers must be 16-byte aligned");
  if (n > (int64_t(1) << 30))
    return fail(GFX_ERR_UNSUPPORTED, "tcgen05 head (TMA): at most 2^30 nodes per call");
  v8h::Maps maps;
  int rc = tma::make_rows128_map(&maps.in, h, n, v8h::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.out, out, n, v8h::kTileM);
  if (rc) return rc;
  v8h::Consts c;
  const gfx_host_vectors &hv = m->host;
  for (int i = 0; i < kHidden; ++i) {
    c.ba[i] = hv.ba[i];
    c.bb[i] = hv.bb[i];
  }
  v8h::Args a{};
  a.wa_img = m->wa_img; a.wb_img = m->wb_img;
  a.n = n;
  GFX_CUDA(cudaFuncSetAttribute(v8h::head8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                v8h::Smem::total));
  const int64_t tiles = (n + v8h::kTileM - 1) / v8h::kTileM;
  const int grid = int(tiles < kNumSMs ? tiles : kNumSMs);
  v8h::head8_kernel<<<grid, v8h::kWarps * 32, v8h::Smem::total, st>>>(maps, c, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx
