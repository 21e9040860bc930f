// Integer kernels: microbatch packing (K4), destination-CSR build (K0),
// core-row map.  All HBM-bound byte/index work; results are bit-exact and
// run-to-run deterministic (atomics are used only where the result does not
// depend on their order).
#include "gfx_common.cuh"

namespace gfx {

// ---------------------------------------------------------------------------
// exclusive scan of int32 (three launches: local scan, scan of block sums,
// add offsets).  Tile = 1024 threads x 4 items.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_inclusive_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix,
// *total receives the block sum.  blockDim.x == kScanThreads.
__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = warp_inclusive_scan(v);
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = warp_sums[lane];
    int sinc = warp_inclusive_scan(s);
    warp_sums[lane] = sinc - s;  // exclusive warp offsets
    if (lane == 31) *total = sinc;
  }
  __syncthreads();
  int result = inc - v + warp_sums[warp];
  __syncthreads();
  return result;
}

__global__ void __launch_bounds__(kScanThreads)
scan_local_kernel(const int32_t *__restrict__ gate, const int *__restrict__ in, int *__restrict__ out,
                  int *__restrict__ block_sums, int64_t n) {
  __shared__ int total;
  if (gate != nullptr && *gate == 0) return;     // gated launch (gfx_csr_build_if): nothing to do
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    sum += v[i];
  }
  int prefix = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = prefix;
    prefix += v[i];
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place (any length)
__global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(const int32_t *__restrict__ gate, int *__restrict__ sums, int count,
                 int64_t *__restrict__ grand_total) {
  __shared__ int total;
  if (gate != nullptr && *gate == 0) return;
  int carry = 0;
  for (int base = 0; base < count; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = i < count ? sums[i] : 0;
    int p = block_exclusive_scan(v, &total);
    if (i < count) sums[i] = p + carry;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(const int32_t *__restrict__ gate, int *__restrict__ out, const int *__restrict__ block_sums,
                int64_t n) {
  if (gate != nullptr && *gate == 0) return;
  const int add = block_sums[blockIdx.x];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) out[base + i] += add;
}

int scan_blocks(int64_t n) { return int((n + kScanTile - 1) / kScanTile); }

// out[i] = sum(in[0..i)) for i in [0,n).  block_sums: scan_blocks(n) ints.
int exclusive_scan(const int *in, int *out, int64_t n, int *block_sums,
                   int64_t *grand_total, cudaStream_t st, const int32_t *gate) {
  if (n == 0) return GFX_OK;
  int nb = scan_blocks(n);
  scan_local_kernel<<<nb, kScanThreads, 0, st>>>(gate, in, out, block_sums, n);
  scan_sums_kernel<<<1, kScanThreads, 0, st>>>(gate, block_sums, nb, grand_total);
  scan_add_kernel<<<nb, kScanThreads, 0, st>>>(gate, out, block_sums, n);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

// ---------------------------------------------------------------------------
// K4  microbatch packing.
// Greedy first-fit over monotone prefix arrays: from record s the batch
// extends to the last stop with node_ptr[stop]-node_ptr[s] <= max_nodes and
// edge_ptr[stop]-edge_ptr[s] <= max_edges, and always takes at least one
// record (api.py:217-221: the limit test is skipped for the first record).
// Phase 1 computes that stop for EVERY s in parallel (two binary searches);
// phase 2 follows the chain from 0.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int64_t last_not_above(const int64_t *__restrict__ a, int64_t lo,
                                                  int64_t hi, int64_t limit) {
  // largest i in [lo, hi] with a[i] <= limit (a non-decreasing, a[lo] <= limit)
  while (lo < hi) {
    int64_t mid = lo + (hi - lo + 1) / 2;
    if (a[mid] <= limit) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void pack_next_kernel(const int64_t *__restrict__ node_ptr,
                                 const int64_t *__restrict__ edge_ptr, int64_t B,
                                 int64_t max_nodes, int64_t max_edges,
                                 int64_t *__restrict__ next_stop) {
  int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= B) return;
  int64_t by_nodes = last_not_above(node_ptr, s, B, node_ptr[s] + max_nodes);
  int64_t by_edges = last_not_above(edge_ptr, s, B, edge_ptr[s] + max_edges);
  int64_t stop = by_nodes < by_edges ? by_nodes : by_edges;
  next_stop[s] = stop > s ? stop : s + 1;
}

__global__ void pack_chase_kernel(const int64_t *__restrict__ next_stop, int64_t B,
                                  int64_t *__restrict__ bounds, int64_t *__restrict__ n_bounds) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t count = 0, s = 0;
  bounds[count++] = 0;
  while (s < B) {
    s = next_stop[s];
    bounds[count++] = s;
  }
  *n_bounds = count;
}

// ---------------------------------------------------------------------------
// K0  destination CSR.
//   1 count      deg[dst-base] += 1                       (integer atomics)
//   2 scan       row_ptr = exclusive_scan(deg), row_ptr[N] = E
//   3 scatter    slot = cursor[dst]++ ; eid[row_ptr[dst]+slot] = e   (any order)
//   4 place      within each row rank edges by their edge id, so the row is in
//                original edge order (= stable sort = the reference's
//                index_add_ summation order) whatever order step 3 ran in;
//                write col_src = src-base, col_type.
// Rows longer than kSmallRow are ranked by a whole block (step 4b).
// ---------------------------------------------------------------------------
constexpr int kSmallRow = 32;

// An edge whose endpoints do not both lie in [base, base + N) is DROPPED (it would index past the
// chunk's arrays) and reported through *status; the reference rejects such a shard when it slices
// it (graph.py:328-330 via GraphShard.slice, "edge index outside shard node range").
__device__ __forceinline__ bool edge_in_range(int32_t s, int32_t d, int32_t base, uint32_t N) {
  return uint32_t(s - base) < N && uint32_t(d - base) < N;
}

__global__ void csr_count_kernel(const int32_t *__restrict__ gate, const int32_t *__restrict__ src,
                                 const int32_t *__restrict__ dst, int64_t E, int32_t base, uint32_t N,
                                 int *__restrict__ deg, int32_t *__restrict__ status) {
  if (gate != nullptr && *gate == 0) return;
  bool bad = false;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < E;
       e += int64_t(gridDim.x) * blockDim.x) {
    const int32_t d = dst[e];
    if (edge_in_range(src[e], d, base, N)) atomicAdd(&deg[d - base], 1); else bad = true;
  }
  if (bad) atomicOr(status, GFX_GRAPH_BAD_EDGE);
}

__global__ void csr_scatter_kernel(const int32_t *__restrict__ gate, const int32_t *__restrict__ src,
                                   const int32_t *__restrict__ dst, int64_t E, int32_t base, uint32_t N,
                                   const int32_t *__restrict__ row_ptr, int *__restrict__ cursor,
                                   int32_t *__restrict__ eid) {
  if (gate != nullptr && *gate == 0) return;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < E;
       e += int64_t(gridDim.x) * blockDim.x) {
    if (!edge_in_range(src[e], dst[e], base, N)) continue;
    int d = dst[e] - base;
    int slot = atomicAdd(&cursor[d], 1);
    eid[row_ptr[d] + slot] = int32_t(e);
  }
}

__global__ void csr_place_kernel(const int32_t *__restrict__ gate, const int32_t *__restrict__ src,
                                 const uint8_t *__restrict__ typ, int32_t base,
                                 const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ eid,
                                 int64_t N, int32_t *__restrict__ col_src, uint8_t *__restrict__ col_type,
                                 int32_t *__restrict__ big_rows, int *__restrict__ n_big) {
  if (gate != nullptr && *gate == 0) return;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int s = row_ptr[i], t = row_ptr[i + 1];
    if (t - s > kSmallRow) {
      big_rows[atomicAdd(n_big, 1)] = int32_t(i);  // order of this list is irrelevant
      continue;
    }
    for (int a = s; a < t; ++a) {
      const int ea = eid[a];
      int rank = 0;
      for (int b = s; b < t; ++b) rank += eid[b] < ea;
      col_src[s + rank] = src[ea] - base;
      col_type[s + rank] = typ[ea];
    }
  }
}

__global__ void csr_place_big_kernel(const int32_t *__restrict__ gate, const int32_t *__restrict__ src,
                                     const uint8_t *__restrict__ typ, int32_t base,
                                     const int32_t *__restrict__ row_ptr,
                                     const int32_t *__restrict__ eid,
                                     const int32_t *__restrict__ big_rows,
                                     const int *__restrict__ n_big,
                                     int32_t *__restrict__ col_src,
                                     uint8_t *__restrict__ col_type) {
  if (gate != nullptr && *gate == 0) return;
  const int count = *n_big;
  for (int r = blockIdx.x; r < count; r += gridDim.x) {
    const int i = big_rows[r];
    const int s = row_ptr[i], t = row_ptr[i + 1];
    for (int a = s + threadIdx.x; a < t; a += blockDim.x) {
      const int ea = eid[a];
      int rank = 0;
      for (int b = s; b < t; ++b) rank += eid[b] < ea;
      col_src[s + rank] = src[ea] - base;
      col_type[s + rank] = typ[ea];
    }
  }
}

// ---------------------------------------------------------------------------
// Row descriptors straight from the edge list (no CSR).  Same result as gfx_row_describe on the
// CSR of the same edges, bit for bit:
//   1 per edge   class k of the edge for its destination d (0 prev: type 0 from d-1, 1 next: type 1
//                from d+1, 2 pair: type 2|3, 3 prev2: type 4 from d-2, 4 next2: type 5 from d+2);
//                state[d] |= 1 << k (one atomicOr), position[k][d] = e; an edge of no class, or a
//                second edge of a class, marks the row GENERIC
//   2 per node   a banded row's edges must also APPEAR in the order prev < next < pair < prev2 <
//                next2 (the order of the CSR row, i.e. the summation order); then the descriptor
//                is assembled, the pairing partner read from the pair edge
// needs_csr[0] becomes 1 when any row is GENERIC (the banded kernel then reads the CSR arrays,
// see gfx_csr_build_if); edges that leave the chunk are reported through *status like K0 does.
// Measured on the 20 M-nt bench shard: 1.35 ms per pass against 1.63 for K0 + gfx_row_describe; an
// atomic-free form (two passes over the edges that cross-check plain stores) took 1.75 ms.
// ---------------------------------------------------------------------------
constexpr uint32_t kEdgeStateGeneric = 1u << 8;

// Both kernels are latency chains (load -> atomic with return -> store; state -> position -> source
// of the pair edge), so each thread carries kEdgeUnroll independent edges / nodes: all loads of a
// group are issued before the first atomic, all atomics before the first use of a returned word.
constexpr int kEdgeUnroll = 4;

__global__ void __launch_bounds__(256)
edge_classify_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                     const uint8_t *__restrict__ typ, int64_t E, int32_t base, uint32_t N,
                     uint32_t *__restrict__ state, int32_t *__restrict__ position,
                     int32_t *__restrict__ status) {
  bool bad = false;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t e0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e0 < E;
       e0 += stride * kEdgeUnroll) {
    int32_t s[kEdgeUnroll], d[kEdgeUnroll], t[kEdgeUnroll];
#pragma unroll
    for (int j = 0; j < kEdgeUnroll; ++j) {
      const int64_t e = e0 + j * stride;
      const bool in = e < E;
      s[j] = in ? src[e] - base : 0;
      d[j] = in ? dst[e] - base : -1;              // -1: no edge here (not reported as bad)
      t[j] = in ? int(typ[e]) : 0;
    }
    uint32_t bit[kEdgeUnroll], old[kEdgeUnroll];
#pragma unroll
    for (int j = 0; j < kEdgeUnroll; ++j) {
      const bool in = e0 + j * stride < E;
      const bool ok = uint32_t(s[j]) < N && uint32_t(d[j]) < N;
      bad |= in && !ok;                            // dropped, as in csr_count_kernel
      const int off = s[j] - d[j];
      int k = -1;
      if (t[j] == 0 && off == -1) k = 0;
      else if (t[j] == 1 && off == 1) k = 1;
      else if (t[j] == 2 || t[j] == 3) k = 2;
      else if (t[j] == 4 && off == -2) k = 3;
      else if (t[j] == 5 && off == 2) k = 4;
      bit[j] = !ok ? 0u : (k < 0 ? kEdgeStateGeneric : (0x10000u | (1u << k)));   // bit 16: classified
    }
#pragma unroll
    for (int j = 0; j < kEdgeUnroll; ++j)
      old[j] = bit[j] ? atomicOr(&state[d[j]], bit[j] & 0xffffu) : 0u;
#pragma unroll
    for (int j = 0; j < kEdgeUnroll; ++j) {
      if (!(bit[j] & 0x10000u)) continue;
      const uint32_t mine = bit[j] & 0xffu;
      if (old[j] & mine) atomicOr(&state[d[j]], kEdgeStateGeneric);   // a second edge of this class
      else position[size_t(__ffs(int(mine)) - 1) * N + d[j]] = int32_t(e0 + j * stride);
    }
  }
  if (bad) atomicOr(status, GFX_GRAPH_BAD_EDGE);
}

__global__ void __launch_bounds__(256)
edge_describe_kernel(const int32_t *__restrict__ src, const uint8_t *__restrict__ typ, int32_t base,
                     int64_t N, const uint32_t *__restrict__ state,
                     const int32_t *__restrict__ position, uint32_t *__restrict__ desc,
                     int32_t *__restrict__ needs_csr) {
  using namespace rowdesc;
  bool generic_seen = false;
  constexpr int U = 2;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 < N; i0 += stride * U) {
    uint32_t st[U];
    int32_t pos[U][5], partner[U], ptyp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) st[u] = i0 + u * stride < N ? state[i0 + u * stride] : 0u;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 5; ++k)                  // a class bit is set by the thread that then
        pos[u][k] = (st[u] >> k) & 1u              // stored the position: valid whenever set
                        ? position[size_t(k) * N + i0 + u * stride] : -1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool pair = (st[u] >> 2) & 1u;
      partner[u] = pair ? src[pos[u][2]] - base : 0;
      ptyp[u] = pair ? int(typ[pos[u][2]]) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= N) continue;
      uint32_t d = 0;
      bool generic = (st[u] & kEdgeStateGeneric) != 0;
      int32_t prev_pos = -1;
      constexpr uint32_t bits[5] = {kPrev, kNext, kPair, kPrev2, kNext2};
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        if (!((st[u] >> k) & 1u)) continue;
        generic |= pos[u][k] < prev_pos;           // not the CSR order of a banded row
        prev_pos = pos[u][k];
        d |= bits[k];
      }
      if ((st[u] >> 2) & 1u) {
        generic |= uint32_t(partner[u]) > kPartnerMask;
        d |= (ptyp[u] == 3 ? kPairRev : 0u) | (uint32_t(partner[u]) << kPartnerShift);
      }
      desc[i0 + u * stride] = generic ? kGeneric : d;
      generic_seen |= generic;
    }
  }
  if (generic_seen) *needs_csr = 1;                // benign race: every writer stores 1
}

// ---------------------------------------------------------------------------
// core-row map
// ---------------------------------------------------------------------------
__global__ void core_flag_kernel(const uint8_t *__restrict__ roles, int64_t N,
                                 int *__restrict__ flag) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += int64_t(gridDim.x) * blockDim.x)
    flag[i] = roles[i] == 0;
}

__global__ void core_map_kernel(const uint8_t *__restrict__ roles, int64_t N,
                                int32_t *__restrict__ out_row) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += int64_t(gridDim.x) * blockDim.x)
    if (roles[i] != 0) out_row[i] = -1;
}

static inline int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = int64_t(kNumSMs) * 16;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}


}  // namespace gfx

using namespace gfx;

extern "C" int gfx_pack_microbatches(const int64_t *node_ptr, const int64_t *edge_ptr, int64_t B,
                                     int64_t max_nodes, int64_t max_edges, int64_t *next_stop,
                                     int64_t *bounds, int64_t *n_bounds, void *stream) {
  if (B <= 0) return fail(GFX_ERR_ARGUMENT, "gfx_pack_microbatches: a shard cannot be empty");
  if (max_nodes <= 0 || max_edges <= 0)
    return fail(GFX_ERR_ARGUMENT, "batch node and edge limits must be positive");
  cudaStream_t st = as_stream(stream);
  StageScope scope(GFX_STAGE_PACK, st, 2);
  pack_next_kernel<<<int((B + 255) / 256), 256, 0, st>>>(node_ptr, edge_ptr, B, max_nodes,
                                                         max_edges, next_stop);
  pack_chase_kernel<<<1, 32, 0, st>>>(next_stop, B, bounds, n_bounds);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

// workspace layout: deg/cursor int[N+1] | eid int[E] | block_sums | big_rows int[N] | n_big, status
extern "C" size_t gfx_csr_workspace_bytes(int64_t N, int64_t E) {
  return align256(size_t(N + 1) * 4) + align256(size_t(E) * 4) +
         align256(size_t(scan_blocks(N + 1) + 1) * 4) + align256(size_t(N) * 4) + 256;
}

extern "C" int gfx_csr_build(const int32_t *edge_src, const int32_t *edge_dst,
                             const uint8_t *edge_type, int64_t N, int64_t E, int32_t node_base,
                             int32_t *row_ptr, int32_t *col_src, uint8_t *col_type, void *ws,
                             size_t ws_bytes, void *stream) {
  return gfx_csr_build_checked(edge_src, edge_dst, edge_type, N, E, node_base, row_ptr, col_src,
                               col_type, nullptr, ws, ws_bytes, stream);
}

static int csr_build(const int32_t *edge_src, const int32_t *edge_dst, const uint8_t *edge_type,
                     int64_t N, int64_t E, int32_t node_base, int32_t *row_ptr, int32_t *col_src,
                     uint8_t *col_type, int32_t *status, const int32_t *gate, void *ws, size_t ws_bytes,
                     void *stream);

extern "C" int gfx_csr_build_checked(const int32_t *edge_src, const int32_t *edge_dst,
                                     const uint8_t *edge_type, int64_t N, int64_t E,
                                     int32_t node_base, int32_t *row_ptr, int32_t *col_src,
                                     uint8_t *col_type, int32_t *status, void *ws,
                                     size_t ws_bytes, void *stream) {
  return csr_build(edge_src, edge_dst, edge_type, N, E, node_base, row_ptr, col_src, col_type, status,
                   nullptr, ws, ws_bytes, stream);
}

extern "C" int gfx_csr_build_if(const int32_t *edge_src, const int32_t *edge_dst,
                                const uint8_t *edge_type, int64_t N, int64_t E, int32_t node_base,
                                int32_t *row_ptr, int32_t *col_src, uint8_t *col_type,
                                const int32_t *needs_csr, void *ws, size_t ws_bytes, void *stream) {
  if (!needs_csr) return fail(GFX_ERR_ARGUMENT, "gfx_csr_build_if: null flag");
  return csr_build(edge_src, edge_dst, edge_type, N, E, node_base, row_ptr, col_src, col_type, nullptr,
                   needs_csr, ws, ws_bytes, stream);
}

static int csr_build(const int32_t *edge_src, const int32_t *edge_dst, const uint8_t *edge_type,
                     int64_t N, int64_t E, int32_t node_base, int32_t *row_ptr, int32_t *col_src,
                     uint8_t *col_type, int32_t *status, const int32_t *gate, void *ws, size_t ws_bytes,
                     void *stream) {
  if (N < 0 || E < 0 || N >= (int64_t(1) << 31) - 1 || E >= (int64_t(1) << 31) - 1)
    return fail(GFX_ERR_ARGUMENT, "gfx_csr_build: sizes must fit int32");
  if (ws_bytes < gfx_csr_workspace_bytes(N, E))
    return fail(GFX_ERR_WORKSPACE, "gfx_csr_build: workspace too small");
  cudaStream_t st = as_stream(stream);
  char *p = static_cast<char *>(ws);
  int *deg = reinterpret_cast<int *>(p); p += align256(size_t(N + 1) * 4);
  int32_t *eid = reinterpret_cast<int32_t *>(p); p += align256(size_t(E) * 4);
  int *sums = reinterpret_cast<int *>(p); p += align256(size_t(scan_blocks(N + 1) + 1) * 4);
  int32_t *big_rows = reinterpret_cast<int32_t *>(p); p += align256(size_t(N) * 4);
  int *n_big = reinterpret_cast<int *>(p);
  if (status == nullptr) status = n_big + 1;     // unobserved: bad edges are still dropped
  StageScope scope(GFX_STAGE_CSR, st, (E > 0 ? 1 : 0) + 3 + ((N == 0 || E == 0) ? 0 : 3));
  GFX_CUDA(cudaMemsetAsync(deg, 0, size_t(N + 1) * 4, st));
  GFX_CUDA(cudaMemsetAsync(n_big, 0, 8, st));
  if (E > 0)
    csr_count_kernel<<<grid_for(E, 256), 256, 0, st>>>(gate, edge_src, edge_dst, E, node_base,
                                                       uint32_t(N), deg, status);
  int rc = exclusive_scan(deg, row_ptr, N + 1, sums, nullptr, st, gate);
  if (rc) return rc;
  if (N == 0 || E == 0) {
    GFX_LAUNCH_CHECK();
    return GFX_OK;
  }
  GFX_CUDA(cudaMemsetAsync(deg, 0, size_t(N + 1) * 4, st));  // reuse as cursor
  csr_scatter_kernel<<<grid_for(E, 256), 256, 0, st>>>(gate, edge_src, edge_dst, E, node_base,
                                                       uint32_t(N), row_ptr, deg, eid);
  csr_place_kernel<<<grid_for(N, 256), 256, 0, st>>>(gate, edge_src, edge_type, node_base, row_ptr,
                                                     eid, N, col_src, col_type, big_rows, n_big);
  csr_place_big_kernel<<<kNumSMs, 256, 0, st>>>(gate, edge_src, edge_type, node_base, row_ptr, eid,
                                                big_rows, n_big, col_src, col_type);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

// workspace: needs_csr int32 (256 B) | state uint32[N] | position int32[5][N]
extern "C" size_t gfx_edge_describe_workspace_bytes(int64_t N) {
  return 256 + align256(size_t(N) * 4) + align256(size_t(N) * 20);
}

extern "C" int gfx_edge_describe(const int32_t *edge_src, const int32_t *edge_dst,
                                 const uint8_t *edge_type, int64_t N, int64_t E, int32_t node_base,
                                 uint32_t *desc, int32_t *status, void *ws, size_t ws_bytes,
                                 void *stream) {
  if (N < 0 || E < 0 || N >= (int64_t(1) << 31) - 1 || E >= (int64_t(1) << 31) - 1)
    return fail(GFX_ERR_ARGUMENT, "gfx_edge_describe: sizes must fit int32");
  if (!ws || ws_bytes < gfx_edge_describe_workspace_bytes(N))
    return fail(GFX_ERR_WORKSPACE, "gfx_edge_describe: workspace too small");
  if (N > 0 && (!desc || !status)) return fail(GFX_ERR_ARGUMENT, "gfx_edge_describe: null array");
  cudaStream_t st = as_stream(stream);
  char *p = static_cast<char *>(ws);
  int32_t *needs_csr = reinterpret_cast<int32_t *>(p); p += 256;
  uint32_t *state = reinterpret_cast<uint32_t *>(p); p += align256(size_t(N) * 4);
  int32_t *position = reinterpret_cast<int32_t *>(p);
  StageScope scope(GFX_STAGE_CSR, st, (E > 0 ? 1 : 0) + (N > 0 ? 1 : 0));
  GFX_CUDA(cudaMemsetAsync(needs_csr, 0, 256, st));
  if (N == 0) return GFX_OK;
  GFX_CUDA(cudaMemsetAsync(state, 0, size_t(N) * 4, st));
  if (E > 0)
    edge_classify_kernel<<<grid_for((E + kEdgeUnroll - 1) / kEdgeUnroll, 256), 256, 0, st>>>(edge_src, edge_dst, edge_type, E, node_base,
                                                           uint32_t(N), state, position, status);
  edge_describe_kernel<<<grid_for((N + 1) / 2, 256), 256, 0, st>>>(edge_src, edge_type, node_base, N, state,
                                                         position, desc, needs_csr);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

extern "C" size_t gfx_core_rows_workspace_bytes(int64_t N) {
  return align256(size_t(N) * 4) + align256(size_t(scan_blocks(N) + 1) * 4);
}

extern "C" int gfx_core_rows(const uint8_t *roles, int64_t N, int32_t *out_row, int64_t *n_core,
                             void *ws, size_t ws_bytes, void *stream) {
  if (N <= 0 || N >= (int64_t(1) << 31) - 1)
    return fail(GFX_ERR_ARGUMENT, "gfx_core_rows: node count must be in [1, 2^31)");
  if (ws_bytes < gfx_core_rows_workspace_bytes(N))
    return fail(GFX_ERR_WORKSPACE, "gfx_core_rows: workspace too small");
  cudaStream_t st = as_stream(stream);
  char *p = static_cast<char *>(ws);
  int *flag = reinterpret_cast<int *>(p); p += align256(size_t(N) * 4);
  int *sums = reinterpret_cast<int *>(p);
  StageScope scope(GFX_STAGE_CORE_ROWS, st, 5);
  core_flag_kernel<<<grid_for(N, 256), 256, 0, st>>>(roles, N, flag);
  int rc = exclusive_scan(flag, out_row, N, sums, n_core, st);
  if (rc) return rc;
  core_map_kernel<<<grid_for(N, 256), 256, 0, st>>>(roles, N, out_row);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}
