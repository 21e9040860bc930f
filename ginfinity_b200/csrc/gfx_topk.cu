// K5  similarity search: per-query top-k of a [Q,128] fp16 query set against a
// [D,128] fp16 embedding database (cosine = dot product of unit vectors, or
// negative squared L2 distance).  No reference counterpart (north-star item 4).
//
//   scan    tcgen05 GEMM  S = Q_tile(256) x DB_tile(128)^T  with the running
//           top-k fused into the TMEM epilogue: the score matrix never exists
//           in memory.  One persistent CTA per SM; a work item is (query tile,
//           database segment).  Warps: 0-7 epilogue (one query per thread),
//           8 MMA issuer, 9 TMA loader.  Database tiles stream through a
//           3-stage TMA ring (the same tile is read by every resident CTA at
//           about the same time, so all but the first read hit L2); the 256
//           queries of the item stay in shared memory; accumulators are
//           double-buffered in TMEM (2 x 2 x 128 columns).
//   finish  exact re-scoring of the surviving candidates (sequential fp32 FMA
//           in dimension order: reproducible and restatable on the CPU) and
//           selection of the k best by (score desc, index asc).
//   merge   k-way merge of per-rank lists after the all-gather.
//
// Tensor-core scores only choose the candidates (kc = k + >=4 spares per
// segment); the returned scores and order come from the exact pass.
#include <limits.h>
#include <math.h>

#include <cstdlib>

#include "gfx_common.cuh"
#include "gfx_tma.cuh"
#include "gfx_umma.cuh"

namespace gfx {
namespace topk {

using namespace ptx;

constexpr int kQTile = 256, kDbTile = 128, kStages = 3;
constexpr int kBoxBytes = 128 * 128;              // [128 rows x 64 cols] fp16
constexpr int kQBytes = 4 * kBoxBytes;            // 2 row blocks x 2 K blocks
constexpr int kDbBytes = 2 * kBoxBytes;
constexpr int kEpiWarps = 8, kMmaWarp = 8, kLoadWarp = 9, kWarps = 10;
constexpr int kD2Slots = 8;                       // loader runs <= 5 tiles ahead of the epilogue
constexpr int kMaxCand = 28, kMaxK = 24;
constexpr int kWaves = 8;                         // work items per CTA aimed for
constexpr int kSampleTiles = 128;                 // database tiles of the threshold-seeding pre-pass

enum Bar {
  kBarQFull = 0, kBarQEmpty = 1, kBarBFull = 2, kBarBEmpty = 5, kBarDFull = 8, kBarDEmpty = 10,
  kNumBars = 12
};

struct Smem {
  static constexpr int off_q = 0;
  static constexpr int off_db = off_q + kQBytes;
  static constexpr int off_d2 = off_db + kStages * kDbBytes;
  static constexpr int off_bar = off_d2 + kD2Slots * kDbTile * 4;
  static constexpr int off_tmem = off_bar + kNumBars * 8;
  static constexpr int off_list = (off_tmem + 8 + 15) & ~15;
  static constexpr int total(int kc) { return off_list + kc * kQTile * 8; }   // per-thread sorted lists
};
static_assert(Smem::total(kMaxCand) <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct alignas(64) Maps {
  CUtensorMap q, db;
};

struct Plan {
  int64_t tiles_q, tiles_db, segs, tiles_per_seg, qp;
  int kc;
  int64_t sample_tiles;       // 0 = no seeding pre-pass
  size_t off_scores, off_index, off_d2, off_seed_scores, off_seed_index, total;
};

static Plan make_plan(int64_t Q, int64_t D, int k) {
  Plan p{};
  p.tiles_q = (Q + kQTile - 1) / kQTile;
  p.tiles_db = (D + kDbTile - 1) / kDbTile;
  p.qp = p.tiles_q * kQTile;
  int64_t segs = 1;
  if (p.tiles_q > 0 && p.tiles_q < int64_t(kWaves) * kNumSMs)
    segs = (int64_t(kWaves) * kNumSMs + p.tiles_q - 1) / p.tiles_q;
  const int64_t max_segs = p.tiles_db / 16 > 1 ? p.tiles_db / 16 : 1;   // >= 16 tiles per segment
  if (segs > max_segs) segs = max_segs;
  p.tiles_per_seg = p.tiles_db > 0 ? (p.tiles_db + segs - 1) / segs : 1;
  p.segs = p.tiles_db > 0 ? (p.tiles_db + p.tiles_per_seg - 1) / p.tiles_per_seg : 1;
  int kc = (k + 4 + 3) & ~3;
  p.kc = kc > kMaxCand ? kMaxCand : kc;
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t cand = size_t(p.segs) * p.qp * p.kc;
  p.off_scores = 0;
  p.off_index = up(cand * 4);
  p.off_d2 = p.off_index + up(cand * 4);
  p.total = p.off_d2 + up(size_t(p.tiles_db) * kDbTile * 4);
  // Seeding pays when the sample is a small part of the scan.
  p.sample_tiles = p.tiles_db >= 8 * kSampleTiles ? kSampleTiles : 0;
  p.off_seed_scores = p.total;
  p.off_seed_index = p.off_seed_scores + up(size_t(p.qp) * p.kc * 4);
  if (p.sample_tiles) p.total = p.off_seed_index + up(size_t(p.qp) * p.kc * 4);
  return p;
}

struct Args {
  const float *d2;            // [tiles_db * 128] squared row norms, +inf past the end (L2 only)
  float *part_scores;         // [segs][qp][kc]
  int32_t *part_index;
  int64_t tiles_q, tiles_db, segs, tiles_per_seg, qp, num_rows;
  int kc;
  // [qp][kc] lists of a pre-pass over the first kSampleTiles tiles, or NULL.  The
  // kc-th best score of ANY subset of the database is a lower bound of the
  // kc-th best overall, so a scan may start from it instead of -inf: it skips
  // the warm-up in which every value is inserted.
  const float *seed;
};

// Insert (v, idx) into the thread's descending list (column `ls`/`li`, stride
// kQTile).  Strict comparisons: among equal scores the earlier (lower) index
// stays ahead.  Returns the new threshold (the list's last score).  (An
// unsorted set with a tracked eviction slot was tried: its full rescan per
// insert costs more than the average shift here.)
__device__ __noinline__ float list_insert(float *ls, int32_t *li, int kc, float v, int32_t idx) {
  int j = kc - 1;
  while (j > 0 && ls[(j - 1) * kQTile] < v) {
    ls[j * kQTile] = ls[(j - 1) * kQTile];
    li[j * kQTile] = li[(j - 1) * kQTile];
    --j;
  }
  ls[j * kQTile] = v;
  li[j * kQTile] = idx;
  return ls[(kc - 1) * kQTile];
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int METRIC>
__global__ void __launch_bounds__(kWarps * 32, 1)
topk_scan_kernel(const __grid_constant__ Maps maps, const Args p) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *qs = smem + L::off_q, *dbs = smem + L::off_db;
  float *d2s = reinterpret_cast<float *>(smem + L::off_d2);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::off_bar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);
  float *list_s = reinterpret_cast<float *>(smem + L::off_list);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
  } else if (tid == 0) {
    mbar_init(bar + kBarQFull, 1);
    mbar_init(bar + kBarQEmpty, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar + kBarBFull + s, 1);
      mbar_init(bar + kBarBEmpty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar + kBarDFull + a, 1);
      mbar_init(bar + kBarDEmpty + a, kEpiWarps);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t items = p.tiles_q * p.segs;

  if (warp == kLoadWarp) {
    // ================================ TMA loader =================================
    // (whole warp in the loop, one elected lane issues: see the MMA issuer below)
    {
      uint32_t elected;
      asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(elected));
      const bool issuer = elected != 0u;
      if (issuer) {
        prefetch_tmap(&maps.q);
        prefetch_tmap(&maps.db);
      }
      uint32_t it = 0, item_n = 0;
      for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
        const int64_t qt = item % p.tiles_q, seg = item / p.tiles_q;
        mbar_wait(bar + kBarQEmpty, (item_n & 1) ^ 1);
        if (issuer) {
          mbar_arrive_expect_tx(bar + kBarQFull, kQBytes);
#pragma unroll
          for (int b = 0; b < 4; ++b)
            tma_load_2d(qs + b * kBoxBytes, &maps.q, (b & 1) * 64, int(qt * kQTile + (b >> 1) * 128),
                        bar + kBarQFull);
        }
        __syncwarp();
        const int64_t t0 = seg * p.tiles_per_seg;
        const int64_t t1 = t0 + p.tiles_per_seg < p.tiles_db ? t0 + p.tiles_per_seg : p.tiles_db;
        for (int64_t t = t0; t < t1; ++t, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          uint8_t *dst = dbs + s * kDbBytes;
          mbar_wait(bar + kBarBEmpty + s, ph ^ 1);
          if (issuer) {
            mbar_arrive_expect_tx(bar + kBarBFull + s, kDbBytes + (METRIC == 1 ? kDbTile * 4 : 0));
            tma_load_2d(dst, &maps.db, 0, int(t * kDbTile), bar + kBarBFull + s);
            tma_load_2d(dst + kBoxBytes, &maps.db, 64, int(t * kDbTile), bar + kBarBFull + s);
            if (METRIC == 1)
              bulk_g2s(d2s + (it % kD2Slots) * kDbTile, p.d2 + t * kDbTile, kDbTile * 4,
                       bar + kBarBFull + s);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer =================================
    // The whole warp runs the loop and waits on the barriers; one elected lane issues.  Issued
    // from inside `if (lane == 0)` the operands of every MMA were ordinary registers to the
    // compiler and each tcgen05.mma came wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop
    // (11-13 instructions per MMA on a warp that shares its scheduler with epilogue warps: more
    // than the 64 cycles an MMA of this shape takes); with warp-uniform operands they live in
    // uniform registers (measured on the fused layer kernel).
    {
      constexpr uint32_t idesc = idesc_f16(128, kDbTile);
      const uint32_t qa = smem_u32(qs), ba = smem_u32(dbs);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      uint32_t elected;
      asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(elected));
      const bool issuer = elected != 0u;
      uint32_t it = 0, item_n = 0;
      for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
        const int64_t seg = item / p.tiles_q;
        const int64_t t0 = seg * p.tiles_per_seg;
        const int64_t t1 = t0 + p.tiles_per_seg < p.tiles_db ? t0 + p.tiles_per_seg : p.tiles_db;
        mbar_wait(bar + kBarQFull, item_n & 1);
        for (int64_t t = t0; t < t1; ++t, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1, a = it & 1, pha = (it >> 1) & 1;
          mbar_wait(bar + kBarBFull + s, ph);
          mbar_wait(bar + kBarDEmpty + a, pha ^ 1);
          tc_fence_after();
          if (issuer) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const int kb = kk >> 2, k = kk & 3;
                const uint64_t da = smem_desc_sw128(qa + (half * 2 + kb) * kBoxBytes + k * 32);
                const uint64_t db = smem_desc_sw128(ba + s * kDbBytes + kb * kBoxBytes + k * 32);
                mma_f16_ss(tm + a * 256 + half * 128, da, db, idesc, kk != 0);
              }
            }
            mma_commit(bar + kBarBEmpty + s);
            mma_commit(bar + kBarDFull + a);
          }
          __syncwarp();
        }
        if (issuer) mma_commit(bar + kBarQEmpty);   // every MMA reading this item's queries is done
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ============ epilogue: running top-k, one query per thread ====================
    const int quad = warp & 3, half = warp >> 2;
    const int row = half * 128 + quad * 32 + lane;          // query inside the item
    const uint32_t tbase = tmem + (uint32_t(quad * 32) << 16) + half * 128;
    const int kc = p.kc;
    const int32_t rows = int32_t(p.num_rows);
    float *ls = list_s + tid;
    int32_t *li = reinterpret_cast<int32_t *>(list_s + kc * kQTile) + tid;
    const float ninf = -INFINITY;
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
      const int64_t qt = item % p.tiles_q, seg = item / p.tiles_q;
      const int64_t t0 = seg * p.tiles_per_seg;
      const int64_t t1 = t0 + p.tiles_per_seg < p.tiles_db ? t0 + p.tiles_per_seg : p.tiles_db;
      for (int j = 0; j < kc; ++j) {
        ls[j * kQTile] = ninf;
        li[j * kQTile] = INT_MAX;
      }
      float seed = ninf;
      if (p.seed) {                                 // kc-th best of the sample (lists are sorted)
        seed = p.seed[(size_t(qt) * kQTile + row) * kc + kc - 1];
        seed -= fabsf(seed) * 2.4e-7f + 1e-30f;     // strictly below: ties with the seed still enter
      }
      float thr = seed;
      for (int64_t t = t0; t < t1; ++t, ++it) {
        const uint32_t a = it & 1, pha = (it >> 1) & 1;
        mbar_wait(bar + kBarDFull + a, pha);
        tc_fence_after();
        float v[kDbTile];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) tmem_ld32(tbase + a * 256 + cb * 32, v + cb * 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + kBarDEmpty + a);   // accumulator is in registers now
        const int32_t base = int32_t(t * kDbTile);
        if (METRIC == 1) {
          // -|q-d|^2 = 2 q.d - |d|^2 - |q|^2 ; the last term is constant per query
          const float4 *d2 = reinterpret_cast<const float4 *>(d2s + (it % kD2Slots) * kDbTile);
#pragma unroll
          for (int c4 = 0; c4 < kDbTile / 4; ++c4) {
            const float4 n = d2[c4];
            v[4 * c4] = fmaf(2.f, v[4 * c4], -n.x);
            v[4 * c4 + 1] = fmaf(2.f, v[4 * c4 + 1], -n.y);
            v[4 * c4 + 2] = fmaf(2.f, v[4 * c4 + 2], -n.z);
            v[4 * c4 + 3] = fmaf(2.f, v[4 * c4 + 3], -n.w);
          }
        }
        // hot path: the maximum of the 128 scores as four independent chains of
        // 3-input maxima.  (Rows past the end of the database -- zero-filled by
        // TMA -- may raise it in the last tile; offer() drops them by index.)
        float t4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          t4[q] = fmaxf(v[32 * q], v[32 * q + 1]);
#pragma unroll
          for (int c = 2; c < 32; c += 2) t4[q] = max3(t4[q], v[32 * q + c], v[32 * q + c + 1]);
        }
        const float top = fmaxf(fmaxf(t4[0], t4[1]), fmaxf(t4[2], t4[3]));
        const bool hit = top > thr;
        if (hit) {
          // rare once the list has warmed up: find the values above the threshold
#pragma unroll
          for (int g = 0; g < kDbTile / 8; ++g) {
            const float gm = fmaxf(max3(v[8 * g], v[8 * g + 1], v[8 * g + 2]),
                                   max3(v[8 * g + 3], v[8 * g + 4],
                                        max3(v[8 * g + 5], v[8 * g + 6], v[8 * g + 7])));
            if (gm > thr) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (v[8 * g + j] > thr && base + 8 * g + j < rows)
                  thr = fmaxf(seed, list_insert(ls, li, kc, v[8 * g + j], base + 8 * g + j));
            }
          }
        }
      }
      // hand the segment's candidates to the finish kernel
      const size_t out = (size_t(seg) * p.qp + size_t(qt) * kQTile + row) * kc;
      for (int j = 0; j < kc; ++j) {
        p.part_scores[out + j] = ls[j * kQTile];
        p.part_index[out + j] = li[j * kQTile];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

// ---- |d|^2 per database row (L2 metric); +inf for the padding rows -----------------
__global__ void __launch_bounds__(256)
row_sqnorm_kernel(const __half *__restrict__ db, int64_t rows, int64_t padded, float *__restrict__ out) {
  const int sub = threadIdx.x & 15;
  const int64_t per_pass = int64_t(gridDim.x) * (blockDim.x >> 4);
  for (int64_t i = int64_t(blockIdx.x) * (blockDim.x >> 4) + (threadIdx.x >> 4); i < padded;
       i += per_pass) {
    float s = 0.f;
    if (i < rows) {
      const uint4 raw = *reinterpret_cast<const uint4 *>(db + i * kHidden + sub * 8);
      const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s = fmaf(f.x, f.x, s);
        s = fmaf(f.y, f.y, s);
      }
    }
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d, 16);
    if (sub == 0) out[i] = i < rows ? s : INFINITY;
  }
}

// ---- ordering -----------------------------------------------------------------------
// (score desc, index asc).  `better(a, b)`: a sorts before b.
template <typename I>
__device__ __forceinline__ bool better(float sa, I ia, float sb, I ib) {
  return sa > sb || (sa == sb && ia < ib);
}

template <typename I>
__device__ __forceinline__ void warp_best(float &s, I &i) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float so = __shfl_xor_sync(0xffffffffu, s, d);
    const I io = __shfl_xor_sync(0xffffffffu, i, d);
    if (better<I>(so, io, s, i)) {
      s = so;
      i = io;
    }
  }
}

// Exact score of (query in shared memory as fp32, database row): a sequential
// fp32 FMA chain in dimension order.
template <int METRIC>
__device__ __forceinline__ float exact_score(const float *q, const __half *row) {
  float acc = 0.f;
#pragma unroll 4
  for (int c = 0; c < kHidden / 8; ++c) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(row) + c);
    const __half *h = reinterpret_cast<const __half *>(&raw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = __half2float(h[j]);
      if (METRIC == 0) {
        acc = fmaf(q[c * 8 + j], x, acc);
      } else {
        const float d = q[c * 8 + j] - x;
        acc = fmaf(d, d, acc);
      }
    }
  }
  return METRIC == 0 ? acc : -acc;
}

constexpr int kFinishWarps = 4;

template <int METRIC>
__global__ void __launch_bounds__(kFinishWarps * 32)
topk_finish_kernel(const __half *__restrict__ queries, const __half *__restrict__ db,
                   float *__restrict__ part_scores, const int32_t *__restrict__ part_index,
                   int64_t segs, int64_t qp, int kc, int64_t Q, int k, int64_t index_base,
                   float *__restrict__ out_scores, int64_t *__restrict__ out_index) {
  __shared__ float qsm[kFinishWarps][kHidden];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = int64_t(blockIdx.x) * kFinishWarps + warp;
  if (qi >= Q) return;
  float *q = qsm[warp];
#pragma unroll
  for (int j = 0; j < 4; ++j) q[lane * 4 + j] = __half2float(queries[qi * kHidden + lane * 4 + j]);
  __syncwarp();
  const int64_t C = segs * kc;
  // pass 1: exact scores, written over the tensor-core scores
  for (int64_t c = lane; c < C; c += 32) {
    const size_t off = (size_t(c / kc) * qp + qi) * kc + (c % kc);
    const int32_t idx = part_index[off];
    part_scores[off] = idx == INT_MAX ? -INFINITY : exact_score<METRIC>(q, db + int64_t(idx) * kHidden);
  }
  __syncwarp();
  // pass 2: k selection rounds over the (score desc, index asc) order
  float last_s = INFINITY;
  int32_t last_i = -1;
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY;
    int32_t bi = INT_MAX;
    for (int64_t c = lane; c < C; c += 32) {
      const size_t off = (size_t(c / kc) * qp + qi) * kc + (c % kc);
      const int32_t i = part_index[off];
      const float s = part_scores[off];
      if (i != INT_MAX && better<int32_t>(last_s, last_i, s, i) && better<int32_t>(s, i, bs, bi)) {
        bs = s;
        bi = i;
      }
    }
    warp_best<int32_t>(bs, bi);
    if (lane == 0) {
      out_scores[qi * k + r] = bi == INT_MAX ? -INFINITY : bs;
      out_index[qi * k + r] = bi == INT_MAX ? int64_t(-1) : index_base + bi;
    }
    if (bi == INT_MAX) {             // fewer than k rows exist: pad the tail
      for (int rr = r + 1 + lane; rr < k; rr += 32) {
        out_scores[qi * k + rr] = -INFINITY;
        out_index[qi * k + rr] = -1;
      }
      break;
    }
    last_s = bs;
    last_i = bi;
  }
}

__global__ void __launch_bounds__(kFinishWarps * 32)
topk_merge_kernel(const float *__restrict__ in_scores, const int64_t *__restrict__ in_index, int parts,
                  int64_t Q, int k, float *__restrict__ out_scores, int64_t *__restrict__ out_index) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = int64_t(blockIdx.x) * kFinishWarps + warp;
  if (qi >= Q) return;
  const int C = parts * k;
  const int64_t none = INT64_MAX;
  float last_s = INFINITY;
  int64_t last_i = -1;
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY;
    int64_t bi = none;
    for (int c = lane; c < C; c += 32) {
      const size_t off = (size_t(c / k) * Q + qi) * k + (c % k);
      const int64_t i = in_index[off];
      const float s = in_scores[off];
      if (i >= 0 && better<int64_t>(last_s, last_i, s, i) && better<int64_t>(s, i, bs, bi)) {
        bs = s;
        bi = i;
      }
    }
    warp_best<int64_t>(bs, bi);
    if (lane == 0) {
      out_scores[qi * k + r] = bi == none ? -INFINITY : bs;
      out_index[qi * k + r] = bi == none ? int64_t(-1) : bi;
    }
    if (bi == none) {
      for (int rr = r + 1 + lane; rr < k; rr += 32) {
        out_scores[qi * k + rr] = -INFINITY;
        out_index[qi * k + rr] = -1;
      }
      break;
    }
    last_s = bs;
    last_i = bi;
  }
}

template <int METRIC>
static int launch_scan(const Maps &maps, const Args &a, int64_t items, cudaStream_t st) {
  const int smem = Smem::total(a.kc);
  GFX_CUDA(cudaFuncSetAttribute(topk_scan_kernel<METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                232448));
  const int grid = int(items < kNumSMs ? items : kNumSMs);
  topk_scan_kernel<METRIC><<<grid, kWarps * 32, smem, st>>>(maps, a);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

}  // namespace topk
}  // namespace gfx

using namespace gfx;

extern "C" size_t gfx_topk_workspace_bytes(int64_t num_queries, int64_t num_rows, int k) {
  if (num_queries <= 0 || num_rows < 0 || k <= 0) return 256;
  return topk::make_plan(num_queries, num_rows, k).total;
}

extern "C" int gfx_topk(const void *queries, int64_t num_queries, const void *database,
                        int64_t num_rows, int dim, int k, int metric, int64_t index_base,
                        float *out_scores, int64_t *out_index, void *workspace,
                        size_t workspace_bytes, void *stream) {
  if (dim != kHidden) return fail(GFX_ERR_UNSUPPORTED, "gfx_topk: kernels are specialised for dim = 128");
  if (k < 1 || k > topk::kMaxK) return fail(GFX_ERR_ARGUMENT, "gfx_topk: k must be in [1, 24]");
  if (metric != 0 && metric != 1) return fail(GFX_ERR_ARGUMENT, "gfx_topk: metric must be 0 (cosine) or 1 (L2)");
  if (num_queries < 0 || num_rows < 0 || num_rows > int64_t(INT_MAX) - 256)
    return fail(GFX_ERR_ARGUMENT, "gfx_topk: bad sizes (rows per call must fit int32)");
  if (num_queries == 0) return GFX_OK;
  if (!queries || !out_scores || !out_index || (num_rows > 0 && !database))
    return fail(GFX_ERR_ARGUMENT, "gfx_topk: null pointer");
  if ((reinterpret_cast<uintptr_t>(queries) | reinterpret_cast<uintptr_t>(database)) & 15)
    return fail(GFX_ERR_ARGUMENT, "gfx_topk: queries and database must be 16-byte aligned");
  topk::Plan plan = topk::make_plan(num_queries, num_rows, k);
  if (workspace_bytes < plan.total || !workspace)
    return fail(GFX_ERR_WORKSPACE, "gfx_topk: workspace too small");
  cudaStream_t st = as_stream(stream);
  char *ws = static_cast<char *>(workspace);
  topk::Args a{};
  a.part_scores = reinterpret_cast<float *>(ws + plan.off_scores);
  a.part_index = reinterpret_cast<int32_t *>(ws + plan.off_index);
  a.d2 = reinterpret_cast<float *>(ws + plan.off_d2);
  a.tiles_q = plan.tiles_q; a.tiles_db = plan.tiles_db; a.segs = plan.segs;
  a.tiles_per_seg = plan.tiles_per_seg; a.qp = plan.qp; a.num_rows = num_rows; a.kc = plan.kc;
  const __half *q = static_cast<const __half *>(queries), *db = static_cast<const __half *>(database);
  const int finish_blocks = int((num_queries + topk::kFinishWarps - 1) / topk::kFinishWarps);
  if (num_rows == 0) {
    // empty database: a merge of zero lists writes (-inf, -1) everywhere
    StageScope scope(GFX_STAGE_TOPK, st, 1);
    topk::topk_merge_kernel<<<finish_blocks, topk::kFinishWarps * 32, 0, st>>>(
        nullptr, nullptr, 0, num_queries, k, out_scores, out_index);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
  }
  topk::Maps maps;
  int rc = tma::make_rows128_map(&maps.q, q, num_queries, 128);
  if (!rc) rc = tma::make_rows128_map(&maps.db, db, num_rows, 128);
  if (rc) return rc;
  StageScope scope(GFX_STAGE_TOPK, st, (metric == 1 ? 3 : 2) + (plan.sample_tiles ? 1 : 0));
  const int64_t items = plan.tiles_q * plan.segs;
  topk::Args pre = a;                      // seeding pre-pass over the first rows, one segment
  static const bool seeding = [] {         // developer switch: GFX_TOPK_SEED=0 scans unseeded
    const char *v = getenv("GFX_TOPK_SEED");
    return !(v && v[0] == '0');
  }();
  if (!seeding) plan.sample_tiles = 0;
  if (plan.sample_tiles) {
    pre.part_scores = reinterpret_cast<float *>(ws + plan.off_seed_scores);
    pre.part_index = reinterpret_cast<int32_t *>(ws + plan.off_seed_index);
    pre.tiles_db = pre.tiles_per_seg = plan.sample_tiles;
    pre.segs = 1;
    pre.num_rows = plan.sample_tiles * topk::kDbTile;
    a.seed = pre.part_scores;
  }
  if (metric == 1) {
    const int64_t padded = plan.tiles_db * topk::kDbTile;
    int64_t blocks = (padded + 15) / 16;
    if (blocks > int64_t(kNumSMs) * 8) blocks = int64_t(kNumSMs) * 8;
    topk::row_sqnorm_kernel<<<int(blocks), 256, 0, st>>>(db, num_rows, padded, const_cast<float *>(a.d2));
    GFX_LAUNCH_CHECK();
    if (plan.sample_tiles && (rc = topk::launch_scan<1>(maps, pre, plan.tiles_q, st))) return rc;
    rc = topk::launch_scan<1>(maps, a, items, st);
    if (rc) return rc;
    topk::topk_finish_kernel<1><<<finish_blocks, topk::kFinishWarps * 32, 0, st>>>(
        q, db, a.part_scores, a.part_index, plan.segs, plan.qp, plan.kc, num_queries, k, index_base,
        out_scores, out_index);
  } else {
    if (plan.sample_tiles && (rc = topk::launch_scan<0>(maps, pre, plan.tiles_q, st))) return rc;
    rc = topk::launch_scan<0>(maps, a, items, st);
    if (rc) return rc;
    topk::topk_finish_kernel<0><<<finish_blocks, topk::kFinishWarps * 32, 0, st>>>(
        q, db, a.part_scores, a.part_index, plan.segs, plan.qp, plan.kc, num_queries, k, index_base,
        out_scores, out_index);
  }
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

extern "C" int gfx_topk_merge(const float *in_scores, const int64_t *in_index, int parts,
                              int64_t num_queries, int k, float *out_scores, int64_t *out_index,
                              void *stream) {
  if (parts < 0 || k < 1 || num_queries < 0)
    return fail(GFX_ERR_ARGUMENT, "gfx_topk_merge: bad sizes");
  if (num_queries == 0) return GFX_OK;
  if (!out_scores || !out_index || (parts > 0 && (!in_scores || !in_index)))
    return fail(GFX_ERR_ARGUMENT, "gfx_topk_merge: null pointer");
  cudaStream_t st = as_stream(stream);
  StageScope scope(GFX_STAGE_TOPK, st, 1);
  const int blocks = int((num_queries + topk::kFinishWarps - 1) / topk::kFinishWarps);
  topk::topk_merge_kernel<<<blocks, topk::kFinishWarps * 32, 0, st>>>(in_scores, in_index, parts,
                                                                     num_queries, k, out_scores, out_index);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}
