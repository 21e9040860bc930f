// One GINE layer in one kernel: K1 (aggregation) is the producer of K2's
// GEMM-1 A operand, so z never exists in HBM.
//
//   h_out = h + LayerNorm(W2 relu(W1' z + b1') + b2),
//   z_i   = (1+eps) h_i + sum_{e: dst=i} relu(h[src_e] + table[type_e])
//
// Why: as two kernels the layer moves ~1,300 B per node through HBM (z written,
// z and h read back, h' written) and runs at two thirds of that bound
// (DESIGN.md section 9); fused it moves ~512 B and the tensor pipe (2,048
// cycles per 128-node tile) becomes the floor.
//
// Status (603k-node chunk, B200): 0.243 ms per layer against 0.195 ms for
// K1 + K2, so the pair stays the default and this kernel is selected by the
// `fused` argument of gfx_encode.  Measured break-down: everything but the
// producers 0.144 ms (bound by the life cycle of the two stage buffers, which
// couples the producers to the store of two tiles earlier), producers working
// from shared memory only +0.06 ms, out-of-block neighbour rows +0.04 ms.  The
// way forward is a CTA pair sharing the weights (64 KB per CTA instead of 128),
// which buys a third and fourth stage buffer.
//
// One persistent CTA per SM, 128-node tiles, 32 warps of 64 registers:
//    0-3   epilogue A   D1 (TMEM) -> + b1, ReLU, fp16 -> A2 (TMEM)
//    4-11  epilogue B   D2 (TMEM) -> + b2, LayerNorm, + residual -> stage buffer
//   12-27  producers    half-warp per node: CSR row -> messages -> z row, written
//                       in the UMMA K-major swizzled layout into the A1 stage
//   28     MMA issuer   (also fetches the weights once)
//   29     store        TMA store of finished stage buffers
//   30     residual     TMA load of the tile's h rows into the stage buffer
//   31     h blocks     bulk copies of 32-row blocks of h into a 3-deep ring
//
// Shared memory (227 KB): W1, W2 operand images (128 KB, resident); two A1
// stage buffers of 32 KB whose life is  z tile -> (MMA-1 done) -> residual
// tile -> LayerNorm output in place -> TMA store -> free;  three 8 KB blocks of
// h rows from which the producers take the in-block neighbours (the graphs
// are near-banded: ~80 % of edges stay inside a 32-row block), the rest come
// from global memory; the fp16 edge table; LayerNorm partial sums.
// The block loader also stages the block's slice of row_ptr / col_src /
// col_type (it is contiguous) in shared memory, so a producer's only global
// accesses are the out-of-block neighbour rows, and those are issued for both
// of its rows of a block before either is summed.
#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"

namespace gfx {

using namespace ptx;

namespace v5 {

constexpr int HID = kMlpHidden, H = HID / 2;
constexpr int kTileM = 128;
constexpr int kTileBytes = kTileM * 128;      // [128 x 64] fp16 box
constexpr int kA1Bytes = 2 * kTileBytes;
constexpr int kBlkRows = 32, kBlkBytes = kBlkRows * 256, kBlkBufs = 3, kBlksPerTile = kTileM / kBlkRows;
constexpr int kWin = 5;                       // edges per row whose rows are loaded up front
// the block's slice of the CSR arrays, staged next to its h rows
constexpr int kRpBytes = 144;                 // row_ptr[r0 .. r0+32] (33 ints) rounded to 16 bytes
constexpr int kCapSrc = 176 + 4;              // col_src ints per block (avg 145) incl. alignment slack
constexpr int kCapTyp = 176 + 16;             // col_type bytes per block incl. alignment slack
constexpr int kIdxBytes = kRpBytes + kCapSrc * 4 + kCapTyp + 16;   // + 16 bytes of meta
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kA2Col = 256, kD2Col = 384;
constexpr int kEpiBWarp0 = 4, kProdWarp0 = 12, kProdWarps = 16, kMmaWarp = 28, kStoreWarp = 29,
              kResWarp = 30, kBlkWarp = 31, kWarps = 32;
constexpr int kHalfWarps = kProdWarps * 2;            // one node row per half-warp at a time
constexpr int kRowsPerBlk = kBlkRows / kHalfWarps;    // rows of a 32-row block per half-warp
static_assert(kRowsPerBlk >= 1 && kRowsPerBlk * kHalfWarps == kBlkRows, "block rows must divide evenly");

enum Bar {
  kBarW = 0, kBarHFull = 1, kBarHEmpty = 4, kBarStageFree = 7, kBarA1Full = 9, kBarA1Empty = 11,
  kBarRFull = 13, kBarD1aFull = 15, kBarD1bFull = 16, kBarA2aFull = 17, kBarA2bFull = 18,
  kBarD2Full = 19, kBarD2Empty = 20, kBarOReady = 21, kNumBars = 23
};

struct Smem {
  static constexpr int w_bytes = HID * kHidden * 2;
  static constexpr int off_w1 = 0;
  static constexpr int off_w2 = off_w1 + w_bytes;
  static constexpr int off_a1 = off_w2 + w_bytes;                   // 2 stages x 32 KB
  static constexpr int off_hs = off_a1 + 2 * kA1Bytes;              // 3 x 8 KB blocks of h rows
  static constexpr int off_tab = off_hs + kBlkBufs * kBlkBytes;     // fp16 [16][128]
  static constexpr int off_xs = off_tab + kMaxEdgeDim * kHidden * 2;
  static constexpr int off_vec = off_xs + 2 * kTileM * 8;           // float b2[128], g[128], b[128]
  static constexpr int off_idx = off_vec + 3 * kHidden * 4;         // 3 x CSR slice of a block
  static constexpr int off_bar = off_idx + kBlkBufs * kIdxBytes;
  static constexpr int off_tmem = off_bar + kNumBars * 8;
  static constexpr int total = off_tmem + 8;
};
static_assert(Smem::total <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct alignas(64) Maps {
  CUtensorMap res, out;      // [n, 128] fp16, box 64 x 128, SWIZZLE_128B
};

struct Consts {
  float b1[HID], b2[kHidden], g[kHidden], b[kHidden];
};

struct Args {
  const __half *h;
  const int32_t *row_ptr, *col_src;
  const uint8_t *col_type;
  const __half *table16, *w1_img, *w2_img;
  int64_t n;
  int edge_dim;
  float eps1;
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
__device__ __forceinline__ uint32_t hfma2_relu_add(uint32_t x, uint32_t t) {
  uint32_t r;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0x3c003c00u), "r"(t));
  return r;
}
__device__ __forceinline__ void add_pair(float &a0, float &a1, uint32_t m) {
  asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.f16 %0, lo, %0;\n"
      "add.rn.f32.f16 %1, hi, %1;\n}"
      : "+f"(a0), "+f"(a1)
      : "r"(m));
}
__device__ __forceinline__ void add_message(float *acc, const uint4 &nb, const uint4 &tb) {
  add_pair(acc[0], acc[1], hfma2_relu_add(nb.x, tb.x));
  add_pair(acc[2], acc[3], hfma2_relu_add(nb.y, tb.y));
  add_pair(acc[4], acc[5], hfma2_relu_add(nb.z, tb.z));
  add_pair(acc[6], acc[7], hfma2_relu_add(nb.w, tb.w));
}
__device__ __forceinline__ uint32_t a_chunk_offset(int r, int c16) {
  return uint32_t(c16 >> 3) * kTileBytes + uint32_t(r) * 128 + uint32_t(((c16 & 7) ^ (r & 7)) << 4);
}

// D1[:, HALF*128 .. +128) -> bias + ReLU -> fp16 -> A2, 32 columns at a time
template <int HALF>
__device__ __forceinline__ void epi_a(const Consts &c, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      int lane) {
  constexpr int col0 = HALF * H;
  mbar_wait(bar + (HALF ? kBarD1bFull : kBarD1aFull), ph);
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
    tmem_st16(trow + kA2Col + col0 / 2 + 16 * q, pk);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar + (HALF ? kBarA2bFull : kBarA2aFull));
}

// D2[:, CH*64 .. +64): pass 1 sums, exchange with the other column half,
// pass 2 (TMEM is read again: cheaper than 64 live registers) normalises and
// adds the residual in place in the stage buffer.
// The per-column vectors come from shared memory here (float4 loads): as
// constant-bank operands they overflow the uniform register file in this
// two-pass form and get spilled, and a spill costs an L2 round trip in a
// kernel whose shared memory leaves no L1.
template <int CH>
__device__ __forceinline__ void epi_b(const float *vec, uint32_t trow, uint64_t *bar, uint32_t ph,
                                      uint32_t ph2, int s, int lane, int quad, float2 *xs,
                                      uint8_t *stage) {
  const float4 *b2v = reinterpret_cast<const float4 *>(vec);
  const float4 *gv = reinterpret_cast<const float4 *>(vec + kHidden);
  const float4 *bv = reinterpret_cast<const float4 *>(vec + 2 * kHidden);
  constexpr int col0 = CH * 64;
  const int r = quad * 32 + lane;
  mbar_wait(bar + kBarD2Full, ph);
  tc_fence_after();
  float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float u[32];
    tmem_ld32(trow + kD2Col + col0 + 32 * q, u);
    tmem_ld_wait();
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 bb = b2v[(col0 + 32 * q) / 4 + j4];
      const float t0 = u[4 * j4] + bb.x, t1 = u[4 * j4 + 1] + bb.y;
      const float t2 = u[4 * j4 + 2] + bb.z, t3 = u[4 * j4 + 3] + bb.w;
      s1[0] += t0; s1[1] += t1; s1[0] += t2; s1[1] += t3;
      s2[0] = fmaf(t0, t0, s2[0]); s2[1] = fmaf(t1, t1, s2[1]);
      s2[0] = fmaf(t2, t2, s2[0]); s2[1] = fmaf(t3, t3, s2[1]);
    }
  }
  xs[CH * kTileM + r] = make_float2(s1[0] + s1[1], s2[0] + s2[1]);
  named_bar_sync(1 + quad, 64);
  const float2 other = xs[(CH ^ 1) * kTileM + r];
  named_bar_sync(1 + quad, 64);
  const float mean = (s1[0] + s1[1] + other.x) * (1.f / kHidden);
  const float var = fmaxf((s2[0] + s2[1] + other.y) * (1.f / kHidden) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nm = -mean * rstd;
  mbar_wait(bar + kBarRFull + s, ph2);
  uint8_t *rrow = stage + CH * kTileBytes + r * 128;
  const int rx = r & 7;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float u[32];
    tmem_ld32(trow + kD2Col + col0 + 32 * q, u);
    tmem_ld_wait();
    if (q == 1) {                                     // D2 fully read
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + kBarD2Empty);
    }
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      uint4 *cell = reinterpret_cast<uint4 *>(rrow + (((q * 4 + gi) ^ rx) << 4));
      const uint4 raw = *cell;
      const __half2 *hp = reinterpret_cast<const __half2 *>(&raw);
      float o[8];
#pragma unroll
      for (int w = 0; w < 2; ++w) {                   // four columns per step
        const int j = gi * 8 + 4 * w, c4 = (col0 + 32 * q + j) / 4;
        const float4 bb = b2v[c4], gg = gv[c4], be = bv[c4];
        const float2 ra = __half22float2(hp[2 * w]), rb = __half22float2(hp[2 * w + 1]);
        o[4 * w] = fmaf(fmaf(u[j] + bb.x, rstd, nm), gg.x, ra.x + be.x);
        o[4 * w + 1] = fmaf(fmaf(u[j + 1] + bb.y, rstd, nm), gg.y, ra.y + be.y);
        o[4 * w + 2] = fmaf(fmaf(u[j + 2] + bb.z, rstd, nm), gg.z, rb.x + be.z);
        o[4 * w + 3] = fmaf(fmaf(u[j + 3] + bb.w, rstd, nm), gg.w, rb.y + be.w);
      }
      *cell = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
    }
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar + kBarOReady + s);
}

__global__ void __launch_bounds__(kWarps * 32, 1)
fused_layer_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Consts c, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *w1s = smem + L::off_w1, *w2s = smem + L::off_w2, *a1s = smem + L::off_a1;
  uint8_t *hss = smem + L::off_hs;
  uint4 *tab = reinterpret_cast<uint4 *>(smem + L::off_tab);
  float2 *xs = reinterpret_cast<float2 *>(smem + L::off_xs);
  float *vec = reinterpret_cast<float *>(smem + L::off_vec);
  uint8_t *idxs = smem + L::off_idx;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::off_bar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
  } else if (tid == 0) {
    mbar_init(bar + kBarW, 1);
    for (int s = 0; s < kBlkBufs; ++s) {
      mbar_init(bar + kBarHFull + s, 1);
      mbar_init(bar + kBarHEmpty + s, kProdWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar + kBarStageFree + s, 1);
      mbar_init(bar + kBarA1Full + s, kProdWarps);
      mbar_init(bar + kBarA1Empty + s, 1);
      mbar_init(bar + kBarRFull + s, 1);
      mbar_init(bar + kBarOReady + s, 8);
    }
    mbar_init(bar + kBarD1aFull, 1);
    mbar_init(bar + kBarD1bFull, 1);
    mbar_init(bar + kBarA2aFull, 4);
    mbar_init(bar + kBarA2bFull, 4);
    mbar_init(bar + kBarD2Full, 1);
    mbar_init(bar + kBarD2Empty, 8);
    fence_mbar_init();
  }
  for (int i = tid; i < p.edge_dim * kHidden / 8; i += blockDim.x)
    tab[i] = reinterpret_cast<const uint4 *>(p.table16)[i];
  for (int i = tid; i < kHidden; i += blockDim.x) {
    vec[i] = c.b2[i];
    vec[kHidden + i] = c.g[i];
    vec[2 * kHidden + i] = c.b[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t tiles = (p.n + kTileM - 1) / kTileM;

  if (warp < kEpiBWarp0) {
    // ================= epilogue A =================================================
    const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      epi_a<0>(c, trow, bar, it & 1, lane);
      epi_a<1>(c, trow, bar, it & 1, lane);
    }
  } else if (warp < kProdWarp0) {
    // ================= epilogue B =================================================
    const int quad = warp & 3, ch = (warp - kEpiBWarp0) >> 2;
    const uint32_t trow = tmem + (uint32_t(quad * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      if (ch == 0)
        epi_b<0>(vec, trow, bar, it & 1, (it >> 1) & 1, s, lane, quad, xs, a1s + s * kA1Bytes);
      else
        epi_b<1>(vec, trow, bar, it & 1, (it >> 1) & 1, s, lane, quad, xs, a1s + s * kA1Bytes);
    }
  } else if (warp < kMmaWarp) {
    // ================= producers: aggregation into the A1 stage ====================
    const int ptid = (warp - kProdWarp0) * 32 + lane;
    const int hw = ptid >> 4, sub = ptid & 15;
    const uint4 *hv = reinterpret_cast<const uint4 *>(p.h) + sub;        // row r -> hv[r * 16]
    const uint4 *tv = tab + sub;
    const int n = int(p.n), ntiles = int(tiles), grid = int(gridDim.x);

    struct Row {              // one node row in flight
      uint4 nb[kWin];
      int beg, end;
      uint32_t types;
    };
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += grid, ++it) {
      const int s = it & 1;
      uint8_t *a1 = a1s + s * kA1Bytes;
      mbar_wait(bar + kBarStageFree + s, ((it >> 1) & 1) ^ 1);
#pragma unroll 1
      for (int blk = 0; blk < kBlksPerTile; ++blk) {
        const uint32_t b = it * kBlksPerTile + blk;
        const uint32_t hb = b % kBlkBufs;
        const uint8_t *hblk = hss + hb * kBlkBytes + sub * 16;
        const uint8_t *idx = idxs + hb * kIdxBytes;
        const int blk_row0 = tile * kTileM + blk * kBlkRows;
        mbar_wait(bar + kBarHFull + hb, (b / kBlkBufs) & 1);
        // meta: {staged?, first staged col_src index, first staged col_type index}
        const int4 meta = *reinterpret_cast<const int4 *>(idx + kIdxBytes - 16);
        const int32_t *rp = reinterpret_cast<const int32_t *>(idx);
        const int32_t *cs = reinterpret_cast<const int32_t *>(idx + kRpBytes) - meta.y;
        const uint8_t *ct = idx + kRpBytes + kCapSrc * 4 - meta.z;
        const bool staged = meta.x != 0;

        auto edge = [&](int e, int &src, int &typ) {
          if (staged) {
            src = cs[e];
            typ = ct[e];
          } else {
            src = p.col_src[e];
            typ = p.col_type[e];
          }
        };
        auto neighbour = [&](int src) -> uint4 {
          const uint32_t local = uint32_t(src - blk_row0);
          return local < uint32_t(kBlkRows) ? *reinterpret_cast<const uint4 *>(hblk + local * 256)
                                            : hv[int64_t(src) * 16];
        };
        auto fetch = [&](int lr, Row &r) {          // indices, then the neighbour-row loads
          const int i = blk_row0 + lr;
          r.beg = r.end = 0;
          r.types = 0;
          if (i < n) {
            r.beg = staged ? rp[lr] : p.row_ptr[i];
            r.end = staged ? rp[lr + 1] : p.row_ptr[i + 1];
          }
#pragma unroll
          for (int u = 0; u < kWin; ++u) {
            if (r.beg + u < r.end) {
              int src, typ;
              edge(r.beg + u, src, typ);
              r.types |= uint32_t(typ) << (4 * u);
              r.nb[u] = neighbour(src);
            }
          }
        };
        auto finish = [&](int lr, const Row &r) {   // messages, self term, z row into A1
          uint4 out = make_uint4(0, 0, 0, 0);
          if (blk_row0 + lr < n) {
            float acc[8];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) acc[ch] = 0.f;
#pragma unroll
            for (int u = 0; u < kWin; ++u)
              if (r.beg + u < r.end) add_message(acc, r.nb[u], tv[((r.types >> (4 * u)) & 15u) * 16]);
            for (int e = r.beg + kWin; e < r.end; ++e) {   // rows longer than the window
              int src, typ;
              edge(e, src, typ);
              add_message(acc, neighbour(src), tv[typ * 16]);
            }
            const uint4 self = *reinterpret_cast<const uint4 *>(hblk + lr * 256);
            const __half2 *sh = reinterpret_cast<const __half2 *>(&self);
            uint32_t *o = reinterpret_cast<uint32_t *>(&out);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float2 f = __half22float2(sh[ch]);
              o[ch] = pack2(fmaf(p.eps1, f.x, acc[2 * ch]), fmaf(p.eps1, f.y, acc[2 * ch + 1]));
            }
          }
          *reinterpret_cast<uint4 *>(a1 + a_chunk_offset(blk * kBlkRows + lr, sub)) = out;
        };
        if (kRowsPerBlk == 2) {
          Row r0, r1;
          fetch(hw, r0);
          fetch(hw + kHalfWarps, r1);                // both rows' loads are in flight
          finish(hw, r0);
          finish(hw + kHalfWarps, r1);
        } else {
          Row r0;
          fetch(hw, r0);
          finish(hw, r0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + kBarHEmpty + hb);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + kBarA1Full + s);
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ====================================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar + kBarW, 2 * L::w_bytes);
      for (int off = 0; off < L::w_bytes; off += 16384) {
        bulk_g2s(w1s + off, reinterpret_cast<const uint8_t *>(p.w1_img) + off, 16384, bar + kBarW);
        bulk_g2s(w2s + off, reinterpret_cast<const uint8_t *>(p.w2_img) + off, 16384, bar + kBarW);
      }
      mbar_wait_parked(bar + kBarW, 0);
      constexpr uint32_t idesc1 = idesc_f16(kTileM, H);
      constexpr uint32_t idesc2 = idesc_f16(kTileM, kHidden);
      const uint32_t w1a = smem_u32(w1s), w2a = smem_u32(w2s);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it & 1, ph2 = (it >> 1) & 1, ph = it & 1;
        const uint32_t a1a = smem_u32(a1s) + s * kA1Bytes;
        mbar_wait_parked(bar + kBarA1Full + s, ph2);
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
        mma_commit(bar + kBarA1Empty + s);      // z has been consumed: the stage can take the residual
        mbar_wait_parked(bar + kBarA2aFull, ph);
        mbar_wait_parked(bar + kBarD2Empty, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HID / 16; ++kk) {
          if (kk == H / 16) {
            mbar_wait_parked(bar + kBarA2bFull, ph);
            tc_fence_after();
          }
          const uint64_t db = smem_desc_sw128(w2a + (kk >> 2) * kTileBytes + (kk & 3) * 32);
          mma_f16_ts(tmem + kD2Col, tmem + kA2Col + kk * 8, db, idesc2, kk != 0);
        }
        mma_commit(bar + kBarD2Full);
      }
    }
    __syncwarp();
  } else if (warp == kStoreWarp) {
    // ============================ output store (TMA) ==============================
    if (lane == 0) {
      prefetch_tmap(&maps.out);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const int row0 = int(tile * kTileM);
        mbar_wait_parked(bar + kBarOReady + s, (it >> 1) & 1);
        tma_store_2d(&maps.out, 0, row0, a1s + s * kA1Bytes);
        tma_store_2d(&maps.out, 64, row0, a1s + s * kA1Bytes + kTileBytes);
        bulk_commit();
        bulk_wait_read<0>();                    // shared memory has been read: the stage is free
        mbar_arrive(bar + kBarStageFree + s);
      }
      bulk_wait_all();
    }
    __syncwarp();
  } else if (warp == kResWarp) {
    // ============================ residual load (TMA) =============================
    if (lane == 0) {
      prefetch_tmap(&maps.res);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const int row0 = int(tile * kTileM);
        mbar_wait_parked(bar + kBarA1Empty + s, (it >> 1) & 1);
        mbar_arrive_expect_tx(bar + kBarRFull + s, kA1Bytes);
        tma_load_2d(a1s + s * kA1Bytes, &maps.res, 0, row0, bar + kBarRFull + s);
        tma_load_2d(a1s + s * kA1Bytes + kTileBytes, &maps.res, 64, row0, bar + kBarRFull + s);
      }
    }
    __syncwarp();
  } else {
    // ============== h rows + CSR slice of every 32-row block (bulk copies) ============
    if (lane == 0) {
      const int n = int(p.n);
      const int num_edges = p.row_ptr[n];
      uint32_t b = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int blk = 0; blk < kBlksPerTile; ++blk, ++b) {
          const uint32_t hb = b % kBlkBufs;
          const int r0 = int(tile) * kTileM + blk * kBlkRows;
          int rows = n - r0;
          rows = rows > kBlkRows ? kBlkRows : (rows < 0 ? 0 : rows);
          uint8_t *idx = idxs + hb * kIdxBytes;
          // the CSR slice: every copy must start and end on 16 bytes inside the arrays
          int e0 = 0, e1 = 0;
          if (rows > 0) {
            e0 = p.row_ptr[r0];
            e1 = p.row_ptr[r0 + rows];
          }
          const int s0 = e0 & ~3, s1 = (e1 + 3) & ~3;          // col_src window (ints)
          const int t0 = e0 & ~15, t1 = (e1 + 15) & ~15;       // col_type window (bytes)
          const bool staged = rows > 0 && r0 + kRpBytes / 4 <= n + 1 && s1 <= num_edges &&
                              t1 <= num_edges && s1 - s0 <= kCapSrc && t1 - t0 <= kCapTyp;
          mbar_wait_parked(bar + kBarHEmpty + hb, ((b / kBlkBufs) & 1) ^ 1);
          *reinterpret_cast<int4 *>(idx + kIdxBytes - 16) = make_int4(staged ? 1 : 0, s0, t0, 0);
          uint32_t bytes = uint32_t(rows * 256);
          if (staged) bytes += kRpBytes + uint32_t(s1 - s0) * 4 + uint32_t(t1 - t0);
          mbar_arrive_expect_tx(bar + kBarHFull + hb, bytes);
          if (rows > 0)
            bulk_g2s(hss + hb * kBlkBytes, p.h + int64_t(r0) * kHidden, uint32_t(rows * 256),
                     bar + kBarHFull + hb);
          if (staged) {
            bulk_g2s(idx, p.row_ptr + r0, kRpBytes, bar + kBarHFull + hb);
            if (s1 > s0)
              bulk_g2s(idx + kRpBytes, p.col_src + s0, uint32_t(s1 - s0) * 4, bar + kBarHFull + hb);
            if (t1 > t0)
              bulk_g2s(idx + kRpBytes + kCapSrc * 4, p.col_type + t0, uint32_t(t1 - t0),
                       bar + kBarHFull + hb);
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace v5

int fused5_layer(const gfx_model *m, int layer, const __half *h, const int32_t *row_ptr,
                 const int32_t *col_src, const uint8_t *col_type, int64_t n, __half *h_out,
                 cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(h_out)) & 15)
    return fail(GFX_ERR_ARGUMENT, "fused layer: activation buffers must be 16-byte aligned");
  if (h == h_out) return fail(GFX_ERR_ARGUMENT, "fused layer: h and h_out must not alias");
  if ((reinterpret_cast<uintptr_t>(row_ptr) | reinterpret_cast<uintptr_t>(col_src) |
       reinterpret_cast<uintptr_t>(col_type)) & 15)
    return fail(GFX_ERR_ARGUMENT,
                "fused layer: row_ptr, col_src and col_type must be 16-byte aligned (their block "
                "slices are fetched with bulk copies)");
  if (n >= (int64_t(1) << 31) - 256) return fail(GFX_ERR_ARGUMENT, "fused layer: too many nodes");
  v5::Maps maps;
  int rc = tma::make_rows128_map(&maps.res, h, n, v5::kTileM);
  if (!rc) rc = tma::make_rows128_map(&maps.out, h_out, n, v5::kTileM);
  if (rc) return rc;
  v5::Consts c;
  const gfx_host_vectors &hv = m->host;
  for (int i = 0; i < kMlpHidden; ++i) c.b1[i] = hv.b1[size_t(layer) * kMlpHidden + i];
  for (int i = 0; i < kHidden; ++i) {
    c.b2[i] = hv.b2[size_t(layer) * kHidden + i];
    c.g[i] = hv.ln_g[size_t(layer) * kHidden + i];
    c.b[i] = hv.ln_b[size_t(layer) * kHidden + i];
  }
  const size_t wi = size_t(layer) * kMlpHidden * kHidden;
  v5::Args a{};
  a.h = h; a.row_ptr = row_ptr; a.col_src = col_src; a.col_type = col_type;
  a.table16 = m->table16 + size_t(layer) * m->edge_dim * kHidden;
  a.w1_img = m->w1_img + wi; a.w2_img = m->w2_img + wi;
  a.n = n; a.edge_dim = m->edge_dim; a.eps1 = m->eps1[layer];
  GFX_CUDA(cudaFuncSetAttribute(v5::fused_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                v5::Smem::total));
  const int64_t tiles = (n + v5::kTileM - 1) / v5::kTileM;
  const int grid = int(tiles < kNumSMs ? tiles : kNumSMs);
  v5::fused_layer_kernel<<<grid, v5::kWarps * 32, v5::Smem::total, st>>>(maps, c, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace gfx
