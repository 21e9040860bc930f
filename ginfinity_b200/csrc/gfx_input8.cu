// Input projection h = x W_in^T + b for fp16 storage, on the tensor cores.
//
//   reference: GINEEncoder.input (nn.Linear(7, 128), _model.py:28,56) on a .half() module: x, W and b
//   in fp16, fp32 accumulation, the result rounded to fp16.
//
// The SIMT kernel (gfx_simt.cu: input_linear4_kernel) issues 56 FMAs per thread for every 8 output
// channels -- 31 warp instructions per node -- and runs at 0.6 of the HBM rate its 256 B per node
// of output would allow (1.4 ms per 20 M-nt pass against 0.9).  Here a 128-row tile is ONE
// tcgen05.mma (M = 128, N = 128, K = 16): the A operand row is (x_0 .. x_6, 1, 0 ...) in fp16, the
// B operand row of output channel c is (W[c][0 .. 6], b[c], 0 ...): the bias rides in the eighth K
// column, as an fp16 value like the reference's.  What is left per node is the epilogue: TMEM ->
// registers, 64 packing conversions, 16 shared-memory stores, and a TMA store of the tile: ~5 warp
// instructions per node.
//
// Two CTAs per SM (96 KB of shared memory, 128 TMEM columns each) overlap one tile's feature
// loads / MMA with the other's epilogue; the output tile is double-buffered against its TMA store.
#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"
#include "gfx_layer_math.cuh"

namespace gfx {

using namespace ptx;

namespace v8i {

using namespace lmath;

constexpr int kTileM = 128;
constexpr int kOpBytes = kTileM * 128;        // an operand tile: [128 rows][128 B], only K slice 0 (32 B) is read
constexpr int kOutBytes = 2 * kTileM * 128;   // [2 K blocks][128 rows][128 B] fp16 output tile
constexpr int kThreads = 128;

struct Smem {
  static constexpr int off_a = 0;
  static constexpr int off_b = off_a + kOpBytes;
  static constexpr int off_out = off_b + kOpBytes;                  // 2 buffers
  static constexpr int off_bar = off_out + 2 * kOutBytes;
  static constexpr int off_tmem = off_bar + 8;
  static constexpr int total = off_tmem + 8;
};

struct alignas(64) Maps {
  CUtensorMap out;           // [n, 128] fp16, box 64 x 128, SWIZZLE_128B
};

__global__ void __launch_bounds__(kThreads, 2)
input_umma_kernel(const __grid_constant__ Maps maps, const float *__restrict__ x,
                  const float *__restrict__ w_in, const float *__restrict__ b_in, int64_t n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Smem::off_bar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Smem::off_tmem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid == 32) {
    mbar_init(bar, 1);
    fence_mbar_init();
    prefetch_tmap(&maps.out);
  }
  // B operand: row c = (W[c][0..6], b[c]) in fp16, then 8 zeros; chunk k (16 B) of row r sits at
  // r * 128 + ((k ^ (r & 7)) << 4) (128-byte swizzle)
  {
    const int c = tid;
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = pack2(w_in[c * kFeat + 2 * j], w_in[c * kFeat + 2 * j + 1]);
    v[3] = pack2(w_in[c * kFeat + 6], b_in[c]);
    const uint32_t row = smem_u32(smem + Smem::off_b) + uint32_t(c) * 128u;
    sts128(row + ((0u ^ uint32_t(c & 7)) << 4), make_uint4(v[0], v[1], v[2], v[3]));
    sts128(row + ((1u ^ uint32_t(c & 7)) << 4), make_uint4(0u, 0u, 0u, 0u));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);

  constexpr uint32_t idesc = idesc_f16(kTileM, kHidden);
  const uint64_t adesc = smem_desc_sw128(smem_u32(smem + Smem::off_a));
  const uint64_t bdesc = smem_desc_sw128(smem_u32(smem + Smem::off_b));
  const int64_t tiles = (n + kTileM - 1) / kTileM;
  const int r = tid;                                   // this thread's tile row = its TMEM lane
  auto features = [&](int64_t tile, float (&f)[kFeat]) {
    const int64_t node = tile * kTileM + r;
#pragma unroll
    for (int j = 0; j < kFeat; ++j) f[j] = (tile < tiles && node < n) ? x[node * kFeat + j] : 0.f;
  };
  float fnext[kFeat];
  features(blockIdx.x, fnext);
  uint32_t it = 0;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    // ---- A operand: this thread's feature row (requested one tile ahead) -----------------------
    float f[kFeat];
#pragma unroll
    for (int j = 0; j < kFeat; ++j) f[j] = fnext[j];
    features(tile + gridDim.x, fnext);
    {
      const uint32_t row = smem_u32(smem + Smem::off_a) + uint32_t(r) * 128u;
      sts128(row + ((0u ^ uint32_t(r & 7)) << 4),
             make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], 1.0f)));
      sts128(row + ((1u ^ uint32_t(r & 7)) << 4), make_uint4(0u, 0u, 0u, 0u));
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_f16_ss(tmem, adesc, bdesc, idesc, 0u);
      mma_commit(bar);
    }
    mbar_wait(bar, it & 1);
    tc_fence_after();
    // ---- epilogue: fp32 accumulators -> fp16 -> swizzled output tile -> TMA store -------------
    const uint32_t out = smem_u32(smem + Smem::off_out) + (it & 1) * kOutBytes;
    if (tid == 0) bulk_wait_read<1>();                 // the store that last read this buffer (two tiles ago)
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[32];
      tmem_ld32(trow + 32 * q, v);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k) {                    // 16-byte chunks 4 q + k of the 256-byte row
        const int c16 = 4 * q + k;
        sts128(out + uint32_t(c16 >> 3) * (kTileM * 128) + sw_off(r, c16 & 7),
               make_uint4(pack2(v[8 * k], v[8 * k + 1]), pack2(v[8 * k + 2], v[8 * k + 3]),
                          pack2(v[8 * k + 4], v[8 * k + 5]), pack2(v[8 * k + 6], v[8 * k + 7])));
      }
    }
    tc_fence_before();
    fence_async_smem();
    __syncthreads();                                   // also: every warp has read its accumulators
    if (tid == 0) {
      const int row0 = int(tile * kTileM);
      tma_store_2d(&maps.out, 0, row0, smem + Smem::off_out + (it & 1) * kOutBytes);
      tma_store_2d(&maps.out, 64, row0, smem + Smem::off_out + (it & 1) * kOutBytes + kTileM * 128);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace v8i

// h must be 16-byte aligned (TMA); n < 2^31 rows.  Returns GFX_ERR_UNSUPPORTED when it cannot run,
// so that the caller falls back to the SIMT kernel.
int input8_linear(const gfx_model *m, const float *x, int64_t n, __half *h, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(h) & 15) || (reinterpret_cast<uintptr_t>(x) & 3) || n >= (int64_t(1) << 31))
    return GFX_ERR_UNSUPPORTED;
  v8i::Maps maps;
  const int rc = tma::make_rows128_map(&maps.out, h, n, v8i::kTileM);
  if (rc) return rc;
  auto kernel = v8i::input_umma_kernel;
  GFX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v8i::Smem::total));
  const int64_t tiles = (n + v8i::kTileM - 1) / v8i::kTileM;
  const int grid = int(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
  kernel<<<grid, v8i::kThreads, v8i::Smem::total, st>>>(maps, x, m->w_in[1], m->b_in, n);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx
