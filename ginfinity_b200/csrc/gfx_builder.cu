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
