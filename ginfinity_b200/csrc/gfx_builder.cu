// K6  graph construction on the device: dot-bracket ->
// pair table -> node features, edges, pointers, for full-molecule records of
// the bundled graph specification (one-hot ACGU, paired flag, sin/cos
// position, backbone + base-pair + skip-2 edges).  Replaces
// GraphBuilder._build_full + _pair_table + GraphShard.from_graphs
// (src/ginfinity/graph.py:494-561, 737-747, 376-412) for that case; every
// array is bit-identical to the reference's.  The two transcendental columns
// cannot be recomputed here bit for bit (NumPy's float32 SIMD sin/cos is not
// correctly rounded), so the host tabulates them once per distinct
// length with NumPy and the fill kernel gathers from that table.
//
//   count   one thread per record walks its structure with an explicit stack
//           (the reference's matcher): partner[], rank of each opening among
//           the record's openings, pairs per record, edges per record
//   scan    edge_ptr = exclusive scan of edges per record (total -> edge_ptr[B])
//   fill    one thread per nucleotide writes its feature row and its (<= 6)
//           edges at their final positions in the reference's edge order
#include "gfx_common.cuh"

namespace gfx {
namespace builder {

enum { kErrBadBase = 1, kErrUnbalanced = 2, kErrBadStructure = 4 };

__global__ void __launch_bounds__(128)
count_kernel(const uint8_t *__restrict__ dbn, const int64_t *__restrict__ node_ptr, int64_t B,
             int skip2, int32_t *__restrict__ partner, int32_t *__restrict__ open_rank,
             int32_t *__restrict__ stack, int32_t *__restrict__ pairs, int32_t *__restrict__ edges,
             int32_t *__restrict__ status) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const int64_t a = node_ptr[r], b = node_ptr[r + 1];
  const int64_t L = b - a;
  int32_t *stk = stack + a;                // the record's own node range is its stack
  int sp = 0, opened = 0, err = 0;
  for (int64_t i = 0; i < L; ++i) {
    const uint8_t ch = dbn[a + i];
    int32_t p = -1;
    if (ch == '(') {
      stk[sp++] = int32_t(i);
      open_rank[a + i] = opened++;
    } else if (ch == ')') {
      if (sp == 0) {
        err |= kErrUnbalanced;
      } else {
        const int32_t j = stk[--sp];
        partner[a + j] = int32_t(i);
        p = j;
      }
    } else if (ch != '.') {
      err |= kErrBadStructure;
    }
    if (ch != '(') partner[a + i] = p;     // openings are written when they close
  }
  if (sp != 0) {
    err |= kErrUnbalanced;
    while (sp > 0) partner[a + stk[--sp]] = -1;
  }
  if (err) atomicOr(status, err);
  pairs[r] = opened;
  const int64_t bb = L > 1 ? L - 1 : 0, sk = (skip2 && L > 2) ? L - 2 : 0;
  edges[r] = int32_t(2 * bb + 2 * int64_t(opened) + 2 * sk);
}

// edge_ptr (int64) from the int32 exclusive scan; edge_ptr[B] = total
__global__ void widen_kernel(const int32_t *__restrict__ scanned, const int64_t *__restrict__ total,
                             int64_t B, int64_t *__restrict__ edge_ptr) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < B) edge_ptr[i] = scanned[i];
  if (i == B) edge_ptr[B] = *total;
}

__device__ __forceinline__ int64_t record_of(const int64_t *__restrict__ node_ptr, int64_t B,
                                             int64_t node) {
  int64_t lo = 0, hi = B - 1;              // largest r with node_ptr[r] <= node
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo + 1) / 2;
    if (node_ptr[mid] <= node) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
fill_kernel(const uint8_t *__restrict__ seq, const uint8_t *__restrict__ dbn,
            const int64_t *__restrict__ node_ptr, const int64_t *__restrict__ edge_ptr, int64_t B,
            int64_t N, int64_t E, int skip2, const int32_t *__restrict__ partner,
            const int32_t *__restrict__ open_rank, const int32_t *__restrict__ pairs,
            const float *__restrict__ pos_table, const int64_t *__restrict__ pos_offset,
            float *__restrict__ feats, int32_t *__restrict__ edge_index,
            uint8_t *__restrict__ edge_types, int32_t *__restrict__ residue_index,
            uint8_t *__restrict__ node_roles, int32_t *__restrict__ status) {
  const int64_t node = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (node >= N) return;
  const int64_t r = record_of(node_ptr, B, node);
  const int64_t g0 = node_ptr[r], L = node_ptr[r + 1] - g0, i = node - g0;
  // ---- feature row: one-hot ACGU | paired | sin, cos -------------------------------
  const uint8_t base = seq[node], ch = dbn[node];
  const int col = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : base == 'U' ? 3 : -1;
  if (col < 0) atomicOr(status, kErrBadBase);
  float *f = feats + node * kFeat;
  f[0] = col == 0 ? 1.f : 0.f;
  f[1] = col == 1 ? 1.f : 0.f;
  f[2] = col == 2 ? 1.f : 0.f;
  f[3] = col == 3 ? 1.f : 0.f;
  f[4] = ch != '.' ? 1.f : 0.f;
  const float2 sc = reinterpret_cast<const float2 *>(pos_table)[pos_offset[r] + i];
  f[5] = sc.x;
  f[6] = sc.y;
  if (residue_index) residue_index[node] = int32_t(i);
  if (node_roles) node_roles[node] = 0;
  // ---- edges, in the reference's order within the record ------------------------------
  int32_t *src = edge_index, *dst = edge_index + E;
  const int64_t e0 = edge_ptr[r];
  const int64_t bb = L > 1 ? L - 1 : 0;
  const int64_t P = pairs[r];
  const int32_t me = int32_t(node);
  if (i < bb) {                              // backbone i -> i+1, then i+1 -> i
    src[e0 + i] = me; dst[e0 + i] = me + 1; edge_types[e0 + i] = 0;
    src[e0 + bb + i] = me + 1; dst[e0 + bb + i] = me; edge_types[e0 + bb + i] = 1;
  }
  if (ch == '(') {                           // pairs by ascending opening, then reversed
    const int32_t j = partner[node];
    if (j >= 0) {
      const int64_t k = open_rank[node];
      const int32_t other = int32_t(g0 + j);
      const int64_t at = e0 + 2 * bb + k;
      src[at] = me; dst[at] = other; edge_types[at] = 2;
      src[at + P] = other; dst[at + P] = me; edge_types[at + P] = 3;
    }
  }
  if (skip2 && i + 2 < L) {                  // skip-2, interleaved
    const int64_t at = e0 + 2 * bb + 2 * P + 2 * i;
    src[at] = me; dst[at] = me + 2; edge_types[at] = 4;
    src[at + 1] = me + 2; dst[at + 1] = me; edge_types[at + 1] = 5;
  }
}

struct Workspace {
  int32_t *partner, *open_rank, *stack, *pairs, *edges, *scanned, *sums;
  int64_t *total;
  int32_t *status;
};

static size_t carve(char *base, int64_t N, int64_t B, Workspace *w) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char *p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  char *p;
  p = take(size_t(N) * 4); if (w) w->partner = reinterpret_cast<int32_t *>(p);
  p = take(size_t(N) * 4); if (w) w->open_rank = reinterpret_cast<int32_t *>(p);
  p = take(size_t(N) * 4); if (w) w->stack = reinterpret_cast<int32_t *>(p);
  p = take(size_t(B) * 4); if (w) w->pairs = reinterpret_cast<int32_t *>(p);
  p = take(size_t(B) * 4); if (w) w->edges = reinterpret_cast<int32_t *>(p);
  p = take(size_t(B) * 4); if (w) w->scanned = reinterpret_cast<int32_t *>(p);
  p = take(size_t(scan_blocks(B) + 1) * 4); if (w) w->sums = reinterpret_cast<int32_t *>(p);
  p = take(8); if (w) w->total = reinterpret_cast<int64_t *>(p);
  p = take(4); if (w) w->status = reinterpret_cast<int32_t *>(p);
  return off;
}


// ---------------------------------------------------------------------------
// K7  windowed ("sliced") records on the device.
// Replaces _select_slice_nodes + _extract_slice + GraphShard.from_graphs
// (src/ginfinity/graph.py:599-695, 376-412): the window [start, end) of a
// molecule is the core; with keep_paired_neighbours, hop 1 adds the pairing
// partners of core nucleotides that lie outside the window and every further
// hop adds all graph neighbours (backbone, pair, skip-2) of the nodes the
// previous hop added; the result is the induced subgraph on the selected
// nodes, nodes in 5'->3' order, edges in the full graph's order, features of
// the FULL molecule (position columns use the full length).
//
//   count        the full-molecule matcher above (partner[] per nucleotide)
//   select       one block per record: window marks, breadth-first hops with a
//                block barrier per hop, then block scans that give every kept
//                node its new index and every kept edge its rank inside its
//                category (backbone / pair / skip-2); per-record totals
//   scan x2      node_ptr and edge_ptr of the sliced shard
//   slice_fill   one thread per nucleotide of the full molecules: kept nodes
//                write their feature row, residue index, role and their kept
//                edges at the final positions
// A record without a window is the window [0, L): everything is kept.
// ---------------------------------------------------------------------------
enum { kErrWindow = 8 };
constexpr int kSelThreads = 256;

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, v, d);
    if (int(threadIdx.x & 31) >= d) v += o;
  }
  return v;
}

// exclusive scan of one int per thread over the block (kSelThreads); *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int *warp_sums, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int inc = warp_incl_scan(v);
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < kSelThreads / 32 ? warp_sums[lane] : 0;
    const int winc = warp_incl_scan(w);
    warp_sums[lane] = winc - w;
    if (lane == 31) warp_sums[32] = winc;
  }
  __syncthreads();
  const int result = inc - v + warp_sums[warp];
  *total = warp_sums[32];
  __syncthreads();
  return result;
}

__global__ void __launch_bounds__(kSelThreads)
slice_select_kernel(const uint8_t *__restrict__ dbn, const int64_t *__restrict__ full_ptr,
                    const int32_t *__restrict__ win_start, const int32_t *__restrict__ win_end,
                    int skip2, int keep_paired, int hops, const int32_t *__restrict__ partner,
                    uint8_t *__restrict__ mark, int32_t *__restrict__ remap,
                    int32_t *__restrict__ pre_bb, int32_t *__restrict__ pre_pair,
                    int32_t *__restrict__ pre_sk, int32_t *__restrict__ n_sel,
                    int32_t *__restrict__ e_sel, int32_t *__restrict__ bb_kept,
                    int32_t *__restrict__ pair_kept, int32_t *__restrict__ status) {
  __shared__ int warp_sums[33];
  const int64_t r = blockIdx.x;
  const int64_t a = full_ptr[r];
  const int L = int(full_ptr[r + 1] - a);
  int s = win_start[r], e = win_end[r];
  if (s < 0 || e > L || s >= e) {                     // the host validates; never trust it
    if (threadIdx.x == 0) atomicOr(status, kErrWindow);
    s = 0;
    e = L;
  }
  const int tid = threadIdx.x;
  uint8_t *mk = mark + a;
  const int32_t *pt = partner + a;
  for (int i = tid; i < L; i += kSelThreads) mk[i] = (i >= s && i < e) ? 1 : 0;
  __syncthreads();
  if (keep_paired) {
    // hop 1: partners of core nucleotides outside the window (graph.py:614-620)
    int any = 0;
    for (int i = s + tid; i < e; i += kSelThreads) {
      const int p = pt[i];
      if (p >= 0 && (p < s || p >= e)) {
        mk[p] = 2;
        any = 1;
      }
    }
    int alive = __syncthreads_or(any);
    // hops 2..: neighbours of the previous hop's additions (graph.py:621-633).
    // marks: 1 = selected earlier, 2 = added by the previous hop, 3 = added by this hop
    for (int hop = 2; hop <= hops && alive; ++hop) {
      any = 0;
      for (int i = tid; i < L; i += kSelThreads) {
        if (mk[i] != 2) continue;
        const int p = pt[i];
        const int cand[5] = {i - 1, i + 1, p, skip2 ? i - 2 : -1, skip2 ? i + 2 : -1};
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const int nb = cand[c];
          if (nb >= 0 && nb < L && mk[nb] == 0) {
            mk[nb] = 3;
            any = 1;
          }
        }
      }
      alive = __syncthreads_or(any);
      for (int i = tid; i < L; i += kSelThreads) {
        const uint8_t m = mk[i];
        if (m >= 2) mk[i] = m - 1;
      }
      __syncthreads();
    }
  }
  // new node indices and per-category edge ranks
  int node_off = 0, bb_off = 0, pair_off = 0, sk_off = 0;
  for (int base = 0; base < L; base += kSelThreads) {
    const int i = base + tid;
    const bool in = i < L;
    const int c = in && mk[i] != 0;
    const int kb = c && i + 1 < L && mk[i + 1] != 0;
    int kp = 0;
    if (c && dbn[a + i] == '(') {
      const int p = pt[i];
      kp = p >= 0 && mk[p] != 0;
    }
    const int ks = skip2 && c && i + 2 < L && mk[i + 2] != 0;
    int t1, t2;
    const int x1 = block_excl_scan(c | (kb << 16), warp_sums, &t1);
    const int x2 = block_excl_scan(kp | (ks << 16), warp_sums, &t2);
    if (in) {
      remap[a + i] = c ? node_off + (x1 & 0xffff) : -1;
      pre_bb[a + i] = bb_off + (x1 >> 16);
      pre_pair[a + i] = pair_off + (x2 & 0xffff);
      pre_sk[a + i] = sk_off + (x2 >> 16);
    }
    node_off += t1 & 0xffff;
    bb_off += t1 >> 16;
    pair_off += t2 & 0xffff;
    sk_off += t2 >> 16;
  }
  if (tid == 0) {
    n_sel[r] = node_off;
    bb_kept[r] = bb_off;
    pair_kept[r] = pair_off;
    e_sel[r] = 2 * (bb_off + pair_off + sk_off);
  }
}

__global__ void __launch_bounds__(256)
slice_fill_kernel(const uint8_t *__restrict__ seq, const uint8_t *__restrict__ dbn,
                  const int64_t *__restrict__ full_ptr, const int32_t *__restrict__ win_start,
                  const int32_t *__restrict__ win_end, const int64_t *__restrict__ node_ptr,
                  const int64_t *__restrict__ edge_ptr, int64_t B, int64_t NF, int64_t E, int skip2,
                  const int32_t *__restrict__ partner, const int32_t *__restrict__ remap,
                  const int32_t *__restrict__ pre_bb, const int32_t *__restrict__ pre_pair,
                  const int32_t *__restrict__ pre_sk, const int32_t *__restrict__ bb_kept,
                  const int32_t *__restrict__ pair_kept, const float *__restrict__ pos_table,
                  const int64_t *__restrict__ pos_offset, float *__restrict__ feats,
                  int32_t *__restrict__ edge_index, uint8_t *__restrict__ edge_types,
                  int32_t *__restrict__ residue_index, uint8_t *__restrict__ node_roles,
                  int32_t *__restrict__ status) {
  const int64_t node = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (node >= NF) return;
  const uint8_t base = seq[node], ch = dbn[node];
  const int col = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : base == 'U' ? 3 : -1;
  if (col < 0) atomicOr(status, kErrBadBase);
  const int32_t m = remap[node];
  if (m < 0) return;
  const int64_t r = record_of(full_ptr, B, node);
  const int64_t a = full_ptr[r], L = full_ptr[r + 1] - a, i = node - a;
  const int64_t g0 = node_ptr[r];
  const int64_t out = g0 + m;
  float *f = feats + out * kFeat;
  f[0] = col == 0 ? 1.f : 0.f;
  f[1] = col == 1 ? 1.f : 0.f;
  f[2] = col == 2 ? 1.f : 0.f;
  f[3] = col == 3 ? 1.f : 0.f;
  f[4] = ch != '.' ? 1.f : 0.f;
  const float2 sc = reinterpret_cast<const float2 *>(pos_table)[pos_offset[r] + i];
  f[5] = sc.x;
  f[6] = sc.y;
  int s = win_start[r], e = win_end[r];
  if (s < 0 || e > L || s >= e) {
    s = 0;
    e = int(L);
  }
  residue_index[out] = int32_t(i);
  node_roles[out] = (i >= s && i < e) ? 0 : 1;
  int32_t *src = edge_index, *dst = edge_index + E;
  const int64_t e0 = edge_ptr[r];
  const int64_t bbK = bb_kept[r], pK = pair_kept[r];
  const int32_t me = int32_t(out);
  if (i + 1 < L) {
    const int32_t nx = remap[node + 1];
    if (nx >= 0) {
      const int64_t k = e0 + pre_bb[node];
      const int32_t other = int32_t(g0 + nx);
      src[k] = me; dst[k] = other; edge_types[k] = 0;
      src[k + bbK] = other; dst[k + bbK] = me; edge_types[k + bbK] = 1;
    }
  }
  if (ch == '(') {
    const int32_t j = partner[node];
    if (j >= 0) {
      const int32_t pj = remap[a + j];
      if (pj >= 0) {
        const int64_t k = e0 + 2 * bbK + pre_pair[node];
        const int32_t other = int32_t(g0 + pj);
        src[k] = me; dst[k] = other; edge_types[k] = 2;
        src[k + pK] = other; dst[k + pK] = me; edge_types[k + pK] = 3;
      }
    }
  }
  if (skip2 && i + 2 < L) {
    const int32_t nx = remap[node + 2];
    if (nx >= 0) {
      const int64_t k = e0 + 2 * bbK + 2 * pK + 2 * int64_t(pre_sk[node]);
      const int32_t other = int32_t(g0 + nx);
      src[k] = me; dst[k] = other; edge_types[k] = 4;
      src[k + 1] = other; dst[k + 1] = me; edge_types[k + 1] = 5;
    }
  }
}

struct SliceWorkspace {
  Workspace full;                    // the matcher's arrays over the full molecules
  uint8_t *mark;
  int32_t *remap, *pre_bb, *pre_pair, *pre_sk;
  int32_t *n_sel, *e_sel, *bb_kept, *pair_kept, *n_scan, *e_scan, *sums;
  int64_t *n_total, *e_total, *full_edge_ptr;
};

static size_t carve_slice(char *base, int64_t NF, int64_t B, SliceWorkspace *w) {
  size_t off = carve(base, NF, B, w ? &w->full : nullptr);
  auto take = [&](size_t bytes) {
    char *p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  char *p;
  p = take(size_t(NF)); if (w) w->mark = reinterpret_cast<uint8_t *>(p);
  p = take(size_t(NF) * 4); if (w) w->remap = reinterpret_cast<int32_t *>(p);
  p = take(size_t(NF) * 4); if (w) w->pre_bb = reinterpret_cast<int32_t *>(p);
  p = take(size_t(NF) * 4); if (w) w->pre_pair = reinterpret_cast<int32_t *>(p);
  p = take(size_t(NF) * 4); if (w) w->pre_sk = reinterpret_cast<int32_t *>(p);
  int32_t **per_record[] = {w ? &w->n_sel : nullptr, w ? &w->e_sel : nullptr,
                            w ? &w->bb_kept : nullptr, w ? &w->pair_kept : nullptr,
                            w ? &w->n_scan : nullptr, w ? &w->e_scan : nullptr};
  for (auto slot : per_record) {
    p = take(size_t(B) * 4);
    if (slot) *slot = reinterpret_cast<int32_t *>(p);
  }
  p = take(size_t(scan_blocks(B) + 1) * 4); if (w) w->sums = reinterpret_cast<int32_t *>(p);
  p = take(8); if (w) w->n_total = reinterpret_cast<int64_t *>(p);
  p = take(8); if (w) w->e_total = reinterpret_cast<int64_t *>(p);
  p = take(size_t(B + 1) * 8); if (w) w->full_edge_ptr = reinterpret_cast<int64_t *>(p);
  return off;
}

}  // namespace builder
}  // namespace gfx

using namespace gfx;

extern "C" size_t gfx_graph_workspace_bytes(int64_t num_nodes, int64_t num_records) {
  if (num_nodes < 0 || num_records < 0) return 0;
  return builder::carve(nullptr, num_nodes, num_records, nullptr);
}

extern "C" int gfx_graph_count(const uint8_t *structures, const int64_t *node_ptr,
                               int64_t num_records, int64_t num_nodes, int skip2,
                               int64_t *edge_ptr, int32_t *status, void *workspace,
                               size_t workspace_bytes, void *stream) {
  if (num_records <= 0) return fail(GFX_ERR_ARGUMENT, "gfx_graph_count: shard is empty");
  if (num_nodes < num_records || num_nodes > int64_t(1) << 30)
    return fail(GFX_ERR_ARGUMENT, "gfx_graph_count: node count must be in [records, 2^30]");
  if (!structures || !node_ptr || !edge_ptr || !status || !workspace)
    return fail(GFX_ERR_ARGUMENT, "gfx_graph_count: null pointer");
  if (workspace_bytes < gfx_graph_workspace_bytes(num_nodes, num_records))
    return fail(GFX_ERR_WORKSPACE, "gfx_graph_count: workspace too small");
  cudaStream_t st = as_stream(stream);
  builder::Workspace w;
  builder::carve(static_cast<char *>(workspace), num_nodes, num_records, &w);
  StageScope scope(GFX_STAGE_BUILD, st, 5);
  GFX_CUDA(cudaMemsetAsync(status, 0, 4, st));
  const int blocks = int((num_records + 127) / 128);
  builder::count_kernel<<<blocks, 128, 0, st>>>(structures, node_ptr, num_records, skip2, w.partner,
                                                w.open_rank, w.stack, w.pairs, w.edges, status);
  GFX_LAUNCH_CHECK();
  int rc = exclusive_scan(w.edges, w.scanned, num_records, w.sums, w.total, st);
  if (rc) return rc;
  const int wb = int((num_records + 1 + 255) / 256);
  builder::widen_kernel<<<wb, 256, 0, st>>>(w.scanned, w.total, num_records, edge_ptr);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

extern "C" int gfx_graph_fill(const uint8_t *sequences, const uint8_t *structures,
                              const int64_t *node_ptr, const int64_t *edge_ptr,
                              int64_t num_records, int64_t num_nodes, int64_t num_edges, int skip2,
                              const float *pos_table, const int64_t *pos_offset,
                              float *node_features, int32_t *edge_index, uint8_t *edge_types,
                              int32_t *residue_index, uint8_t *node_roles, int32_t *status,
                              void *workspace, size_t workspace_bytes, void *stream) {
  if (num_records <= 0 || num_nodes <= 0)
    return fail(GFX_ERR_ARGUMENT, "gfx_graph_fill: shard is empty");
  if (!sequences || !structures || !node_ptr || !edge_ptr || !pos_table || !pos_offset ||
      !node_features || !status || !workspace || (num_edges > 0 && (!edge_index || !edge_types)))
    return fail(GFX_ERR_ARGUMENT, "gfx_graph_fill: null pointer");
  if (workspace_bytes < gfx_graph_workspace_bytes(num_nodes, num_records))
    return fail(GFX_ERR_WORKSPACE, "gfx_graph_fill: workspace too small");
  cudaStream_t st = as_stream(stream);
  builder::Workspace w;
  builder::carve(static_cast<char *>(workspace), num_nodes, num_records, &w);
  StageScope scope(GFX_STAGE_BUILD, st, 1);
  const int blocks = int((num_nodes + 255) / 256);
  builder::fill_kernel<<<blocks, 256, 0, st>>>(
      sequences, structures, node_ptr, edge_ptr, num_records, num_nodes, num_edges, skip2, w.partner,
      w.open_rank, w.pairs, pos_table, pos_offset, node_features, edge_index, edge_types,
      residue_index, node_roles, status);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

extern "C" size_t gfx_slice_workspace_bytes(int64_t num_full_nodes, int64_t num_records) {
  if (num_full_nodes < 0 || num_records < 0) return 0;
  return builder::carve_slice(nullptr, num_full_nodes, num_records, nullptr);
}

extern "C" int gfx_slice_select(const uint8_t *structures, const int64_t *full_ptr,
                                const int32_t *win_start, const int32_t *win_end,
                                int64_t num_records, int64_t num_full_nodes, int skip2,
                                int keep_paired_neighbours, int context_hops, int64_t *node_ptr,
                                int64_t *edge_ptr, int32_t *status, void *workspace,
                                size_t workspace_bytes, void *stream) {
  if (num_records <= 0) return fail(GFX_ERR_ARGUMENT, "gfx_slice_select: shard is empty");
  if (num_full_nodes < num_records || num_full_nodes > int64_t(1) << 30)
    return fail(GFX_ERR_ARGUMENT, "gfx_slice_select: node count must be in [records, 2^30]");
  if (context_hops < 1) return fail(GFX_ERR_ARGUMENT, "gfx_slice_select: context_hops must be >= 1");
  if (!structures || !full_ptr || !win_start || !win_end || !node_ptr || !edge_ptr || !status ||
      !workspace)
    return fail(GFX_ERR_ARGUMENT, "gfx_slice_select: null pointer");
  if (workspace_bytes < gfx_slice_workspace_bytes(num_full_nodes, num_records))
    return fail(GFX_ERR_WORKSPACE, "gfx_slice_select: workspace too small");
  if (num_records > 0x7fffffff) return fail(GFX_ERR_ARGUMENT, "gfx_slice_select: too many records");
  cudaStream_t st = as_stream(stream);
  builder::SliceWorkspace w;
  builder::carve_slice(static_cast<char *>(workspace), num_full_nodes, num_records, &w);
  StageScope scope(GFX_STAGE_BUILD, st, 10);
  GFX_CUDA(cudaMemsetAsync(status, 0, 4, st));
  const int blocks = int((num_records + 127) / 128);
  builder::count_kernel<<<blocks, 128, 0, st>>>(structures, full_ptr, num_records, skip2,
                                                w.full.partner, w.full.open_rank, w.full.stack,
                                                w.full.pairs, w.full.edges, status);
  GFX_LAUNCH_CHECK();
  builder::slice_select_kernel<<<int(num_records), builder::kSelThreads, 0, st>>>(
      structures, full_ptr, win_start, win_end, skip2, keep_paired_neighbours, context_hops,
      w.full.partner, w.mark, w.remap, w.pre_bb, w.pre_pair, w.pre_sk, w.n_sel, w.e_sel, w.bb_kept,
      w.pair_kept, status);
  GFX_LAUNCH_CHECK();
  int rc = exclusive_scan(w.n_sel, w.n_scan, num_records, w.sums, w.n_total, st);
  if (rc) return rc;
  const int wb = int((num_records + 1 + 255) / 256);
  builder::widen_kernel<<<wb, 256, 0, st>>>(w.n_scan, w.n_total, num_records, node_ptr);
  rc = exclusive_scan(w.e_sel, w.e_scan, num_records, w.sums, w.e_total, st);
  if (rc) return rc;
  builder::widen_kernel<<<wb, 256, 0, st>>>(w.e_scan, w.e_total, num_records, edge_ptr);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}

extern "C" int gfx_slice_fill(const uint8_t *sequences, const uint8_t *structures,
                              const int64_t *full_ptr, const int32_t *win_start,
                              const int32_t *win_end, const int64_t *node_ptr,
                              const int64_t *edge_ptr, int64_t num_records, int64_t num_full_nodes,
                              int64_t num_nodes, int64_t num_edges, int skip2,
                              const float *pos_table, const int64_t *pos_offset,
                              float *node_features, int32_t *edge_index, uint8_t *edge_types,
                              int32_t *residue_index, uint8_t *node_roles, int32_t *status,
                              void *workspace, size_t workspace_bytes, void *stream) {
  if (num_records <= 0 || num_full_nodes <= 0 || num_nodes <= 0)
    return fail(GFX_ERR_ARGUMENT, "gfx_slice_fill: shard is empty");
  if (!sequences || !structures || !full_ptr || !win_start || !win_end || !node_ptr || !edge_ptr ||
      !pos_table || !pos_offset || !node_features || !residue_index || !node_roles || !status ||
      !workspace || (num_edges > 0 && (!edge_index || !edge_types)))
    return fail(GFX_ERR_ARGUMENT, "gfx_slice_fill: null pointer");
  if (workspace_bytes < gfx_slice_workspace_bytes(num_full_nodes, num_records))
    return fail(GFX_ERR_WORKSPACE, "gfx_slice_fill: workspace too small");
  cudaStream_t st = as_stream(stream);
  builder::SliceWorkspace w;
  builder::carve_slice(static_cast<char *>(workspace), num_full_nodes, num_records, &w);
  StageScope scope(GFX_STAGE_BUILD, st, 1);
  const int blocks = int((num_full_nodes + 255) / 256);
  builder::slice_fill_kernel<<<blocks, 256, 0, st>>>(
      sequences, structures, full_ptr, win_start, win_end, node_ptr, edge_ptr, num_records,
      num_full_nodes, num_edges, skip2, w.full.partner, w.remap, w.pre_bb, w.pre_pair, w.pre_sk,
      w.bb_kept, w.pair_kept, pos_table, pos_offset, node_features, edge_index, edge_types,
      residue_index, node_roles, status);
  GFX_LAUNCH_CHECK();
  return GFX_OK;
}
