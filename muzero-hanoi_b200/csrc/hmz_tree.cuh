// Device-side building blocks of the tree store, shared by the stand-alone select / backup
// kernels (hmz_tree.cu) and the fused simulation kernels (hmz_fused_*.cu).
#pragma once
#include <math.h>

#include "hmz_common.cuh"

namespace hmz {

static_assert(sizeof(hmz_node_t) == 128, "hmz_node_t must be one 128-byte line");

struct Leaf {
  int parent;  // record of the leaf's parent
  int action;  // action from that parent to the (unexpanded) leaf
  int depth;   // number of best_child steps taken (>= 1)
};

// Writes a freshly expanded node (Node.expand, MCTS/node.py:30-51): six children with priors
// `pr`, N = 0, W = 0, rwd = 0, no expanded grandchildren.  The eight lanes of a segment store
// one 16-byte chunk each, i.e. one coalesced 128-byte line.
__device__ __forceinline__ void write_fresh_record(hmz_node_t* rec, int lane8, const float (&pr)[6], int parent,
                                                   int parent_action) {
  uint4 c = make_uint4(0u, 0u, 0u, 0u);
  if (lane8 == 3) c = make_uint4(__float_as_uint(pr[0]), __float_as_uint(pr[1]), __float_as_uint(pr[2]), __float_as_uint(pr[3]));
  if (lane8 == 4) c = make_uint4(__float_as_uint(pr[4]), __float_as_uint(pr[5]), 0u, 0u);
  if (lane8 == 6) c = make_uint4(0u, 0u, 0u, 0xFFFFFFFFu);
  if (lane8 == 7) c = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, (uint32_t)parent | ((uint32_t)parent_action << 16), 0u);
  reinterpret_cast<uint4*>(rec)[lane8] = c;
}

constexpr int kPathCap = 32;  // path levels kept for the lane-parallel backup (deeper paths walk parent links)

// Node.best_child repeated from the root until an unexpanded child (MCTS/mcts.py:80-86,
// MCTS/node.py:72-123).  Called by the 8 lanes of the warp segment that owns the search; every
// shuffle uses the segment's own mask, so the four segments of a warp run their (different) path
// depths independently and a finished segment issues nothing more.
//   root_n        root.N  (= number of completed simulations)
//   root_prior64  float64 root priors when the root was Dirichlet-noised, else nullptr
//   path_out      nullable [path_cap] bytes: chosen action per level (diagnostics)
//   path_ent      nullable [kPathCap] words: (record | action << 16) per level, for backup_levels()
template <bool kPrefetch = false>
__device__ __forceinline__ Leaf select_leaf(const hmz_node_t* nodes, const double* __restrict__ root_prior64, double mn,
                                            double mx, int root_n, const double* __restrict__ ucb_table,
                                            double discount, int lane8, bool valid, uint8_t* __restrict__ path_out,
                                            int path_cap, uint32_t* __restrict__ path_ent) {
  const int seg_base = (threadIdx.x & 31) & ~7;
  const unsigned seg = 0xFFu << seg_base;
  const bool normalise = mx > mn;  // MinMaxStats.normalize (MCTS/utils_mcts.py:12-16)
  const double range = __dsub_rn(mx, mn);
  int e = 0, n_parent = root_n, depth = 0;
  bool active = valid;
  Leaf leaf{0, 0, 0};
  uint32_t ent0 = 0, ent1 = 0, ent2 = 0, ent3 = 0;  // this lane's path entries for levels lane8 + {0, 8, 16, 24}
  while (active) {
    float score = -INFINITY;
    int c_n = 0, c_child = (int)HMZ_NO_CHILD;
    const hmz_node_t* rec = nodes + e;
    if (lane8 < 6) {
      c_n = rec->N[lane8];
      c_child = rec->child[lane8];
    }
    if (kPrefetch) {  // start fetching the record of the most-visited expanded child while the scores are computed
      int key = (c_child != (int)HMZ_NO_CHILD) ? (((c_n + 1) << 3) | (7 - lane8)) : 0;
#pragma unroll
      for (int off = 4; off; off >>= 1) key = max(key, __shfl_xor_sync(seg, key, off));
      const int likely = __shfl_sync(seg, c_child, seg_base | (7 - (key & 7)));
      if (key != 0 && lane8 == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + likely));
    }
    if (lane8 < 6) {
      const double w_sum = rec->W[lane8];
      const float prior = rec->prior[lane8];
      const float rwd = rec->rwd[lane8];
      float qf = 0.0f;  // child_Q: 0 for unvisited children (node.py:98-102)
      if (c_n > 0) {
        double q = __dadd_rn((double)rwd, __dmul_rn(discount, __ddiv_rn(w_sum, (double)c_n)));
        if (normalise) q = __ddiv_rn(__dsub_rn(q, mn), range);
        qf = __double2float_rn(q);
      }
      // child_U: w = (log((N+c_base+1)/c_base) + c_init) * sqrt(N) / (child.N + 1)  (node.py:114-121)
      const double w = __ddiv_rn(ucb_table[n_parent], (double)(c_n + 1));
      float u;
      if (e == 0 && root_prior64 != nullptr)
        u = __double2float_rn(__dmul_rn(root_prior64[lane8], w));  // float64 prior: product in float64
      else
        u = __fmul_rn(prior, __double2float_rn(w));  // float32 prior: weak scalar -> float32 product
      score = __fadd_rn(qf, u);                     // node.py:83 on float32 arrays
    }
    int best = lane8;
#pragma unroll
    for (int off = 4; off; off >>= 1) {
      const float os = __shfl_xor_sync(seg, score, off);
      const int ob = __shfl_xor_sync(seg, best, off);
      if (os > score || (os == score && ob < best)) {
        score = os;
        best = ob;
      }
    }
    const int b_child = __shfl_sync(seg, c_child, seg_base | best);
    const int b_n = __shfl_sync(seg, c_n, seg_base | best);
    if (path_out != nullptr && lane8 == 0 && depth < path_cap) path_out[depth] = (uint8_t)best;
    if ((depth & 7) == lane8) {
      const uint32_t ent = (uint32_t)e | ((uint32_t)best << 16);
      const int slot = depth >> 3;
      if (slot == 0) ent0 = ent; else if (slot == 1) ent1 = ent; else if (slot == 2) ent2 = ent; else if (slot == 3) ent3 = ent;
    }
    ++depth;
    if (b_child == (int)HMZ_NO_CHILD) {
      leaf.parent = e;
      leaf.action = best;
      leaf.depth = depth;
      active = false;
    } else {
      e = b_child;
      n_parent = b_n;
    }
  }
  if (valid && path_ent != nullptr) {  // 32 contiguous bytes per 8 levels
    path_ent[lane8] = ent0;
    if (leaf.depth > 8) path_ent[8 + lane8] = ent1;
    if (leaf.depth > 16) path_ent[16 + lane8] = ent2;
    if (leaf.depth > 24) path_ent[24 + lane8] = ent3;
  }
  return leaf;
}

__device__ __forceinline__ void minmax_update(double x, double& mn, double& mx) {
  if (x > mx) mx = x;  // python max(maximum, value): value only when strictly greater
  if (x < mn) mn = x;
}

// node.expand bookkeeping on the parent slot + Node.backup (MCTS/node.py:53-70) from the leaf to
// the root, one lane per search.  `value` enters as the network value of the new node.
__device__ __forceinline__ void backup_path(hmz_node_t* __restrict__ nodes, int pe, int pa, int sim, float r,
                                            double value, double discount, double& root_w, double& mn, double& mx) {
  nodes[pe].rwd[pa] = r;                      // leaf.rwd = reward (node.py:44)
  nodes[pe].child[pa] = (uint16_t)(sim + 1);  // leaf is now expanded: its record
  int e = pe, a = pa;
  double rwd = (double)r;
  while (true) {
    hmz_node_t* rec = nodes + e;
    const double w_sum = __dadd_rn(rec->W[a], value);  // current.W += value
    const int n = (int)rec->N[a] + 1;                  // current.N += 1
    rec->W[a] = w_sum;
    rec->N[a] = (uint16_t)n;
    const double q = __ddiv_rn(w_sum, (double)n);
    minmax_update(__dadd_rn(rwd, __dmul_rn(discount, q)), mn, mx);
    value = __dadd_rn(rwd, __dmul_rn(discount, value));  // value = rwd + discount * value
    if (e == 0) break;
    a = rec->parent_action;
    e = rec->parent;
    rwd = (double)nodes[e].rwd[a];
  }
  // the root itself: rwd = 0.0 (MCTS/mcts.py:69), N = sim + 1 after this backup
  root_w = __dadd_rn(root_w, value);
  const double q = __ddiv_rn(root_w, (double)(sim + 1));
  minmax_update(__dadd_rn(0.0, __dmul_rn(discount, q)), mn, mx);
}

// Lane-parallel form of the same backup for paths recorded by select_leaf (depth <= kPathCap):
// the 8 lanes of the segment load their path levels' slots at once (one memory round trip instead
// of `depth` dependent ones), the value recurrence value_k = rwd_{k+1} + discount * value_{k+1}
// runs through segment shuffles in the reference's leaf-to-root order, and each lane then applies
// W += value_k, N += 1 and its min/max candidate.  min/max are order-independent, every float64
// operation is the same as in backup_path, so results are bit-identical.
__device__ __forceinline__ void backup_levels(hmz_node_t* nodes, const uint32_t* __restrict__ path_ent, int depth, int pe,
                                              int pa, int sim, float r, double value, double discount, double& root_w,
                                              double& mn, double& mx, int lane8) {
  const int seg_base = (threadIdx.x & 31) & ~7;
  const unsigned seg = 0xFFu << seg_base;
  double cur = value, lmn = mn, lmx = mx;
  for (int c = (depth - 1) >> 3; c >= 0; --c) {
    const int k = c * 8 + lane8;
    const bool has = k < depth;
    int e = 0, a = 0, n = 0;
    double w_sum = 0.0, rwd = 0.0;
    if (has) {
      const uint32_t ent = path_ent[k];
      e = (int)(ent & 0xFFFFu);
      a = (int)(ent >> 16);
      w_sum = nodes[e].W[a];
      n = nodes[e].N[a];
      rwd = (k == depth - 1) ? (double)r : (double)nodes[e].rwd[a];
    }
    double mine = 0.0;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      const double rj = __shfl_sync(seg, rwd, seg_base | j);
      if (c * 8 + j < depth) {  // segment-uniform
        if (lane8 == j) mine = cur;
        cur = __dadd_rn(rj, __dmul_rn(discount, cur));
      }
    }
    if (has) {
      w_sum = __dadd_rn(w_sum, mine);
      n += 1;
      nodes[e].W[a] = w_sum;
      nodes[e].N[a] = (uint16_t)n;
      if (k == depth - 1) {  // the leaf slot: Node.expand bookkeeping on the parent (node.py:44-49)
        nodes[pe].rwd[pa] = r;
        nodes[pe].child[pa] = (uint16_t)(sim + 1);
      }
      minmax_update(__dadd_rn(rwd, __dmul_rn(discount, __ddiv_rn(w_sum, (double)n))), lmn, lmx);
    }
  }
#pragma unroll
  for (int off = 4; off; off >>= 1) {
    lmn = fmin(lmn, __shfl_xor_sync(seg, lmn, off));
    lmx = fmax(lmx, __shfl_xor_sync(seg, lmx, off));
  }
  root_w = __dadd_rn(root_w, cur);
  minmax_update(__dadd_rn(0.0, __dmul_rn(discount, __ddiv_rn(root_w, (double)(sim + 1)))), lmn, lmx);
  mn = lmn;
  mx = lmx;
}

// ---------------------------------------------------------------------------------------------
// Thread-per-search forms (used by the fused hot loop).  ncu on the 8-lane kernels showed them
// issue-bound (51 % issue-active, ~145 warp instructions per tree level for 4 searches, a quarter of
// the lanes idle, shuffle overhead): one thread per search evaluates the six children with
// instruction-level parallelism instead and cuts warp instructions per search-level about 3x.
// Arithmetic and tie-breaks are operation-for-operation those of select_leaf / backup_path.

struct RecView {  // one 128-byte record in registers
  uint4 q[8];
  __device__ __forceinline__ void load(const hmz_node_t* rec) {
    const uint4* p = reinterpret_cast<const uint4*>(rec);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = p[i];
  }
  __device__ __forceinline__ double W(int a) const {
    const uint4& v = q[a >> 1];
    return (a & 1) ? __hiloint2double((int)v.w, (int)v.z) : __hiloint2double((int)v.y, (int)v.x);
  }
  __device__ __forceinline__ uint32_t word(int i) const {  // 32-bit word i of the record
    const uint4& v = q[i >> 2];
    const int j = i & 3;
    return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
  }
  __device__ __forceinline__ float prior(int a) const { return __uint_as_float(word(12 + a)); }
  __device__ __forceinline__ float rwd(int a) const { return __uint_as_float(word(18 + a)); }
  __device__ __forceinline__ int N(int a) const { return (int)((word(24 + (a >> 1)) >> ((a & 1) * 16)) & 0xFFFFu); }
  __device__ __forceinline__ int child(int a) const { return (int)((word(27 + (a >> 1)) >> ((a & 1) * 16)) & 0xFFFFu); }
};

__device__ __forceinline__ Leaf select_leaf_thread(const hmz_node_t* nodes, const double* __restrict__ root_prior64, double mn,
                                                   double mx, int root_n, const double* __restrict__ ucb_table,
                                                   double discount, uint32_t* __restrict__ path_ent) {
  const bool normalise = mx > mn;
  const double range = __dsub_rn(mx, mn);
  int e = 0, n_parent = root_n, depth = 0;
  Leaf leaf{0, 0, 0};
  while (true) {
    RecView rec;
    rec.load(nodes + e);
    const double tn = ucb_table[n_parent];
    float best_score = 0.f;
    int best = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const int c_n = rec.N(a);
      float qf = 0.0f;
      if (c_n > 0) {
        double q = __dadd_rn((double)rec.rwd(a), __dmul_rn(discount, __ddiv_rn(rec.W(a), (double)c_n)));
        if (normalise) q = __ddiv_rn(__dsub_rn(q, mn), range);
        qf = __double2float_rn(q);
      }
      const double w = __ddiv_rn(tn, (double)(c_n + 1));
      float u;
      if (e == 0 && root_prior64 != nullptr)
        u = __double2float_rn(__dmul_rn(root_prior64[a], w));
      else
        u = __fmul_rn(rec.prior(a), __double2float_rn(w));
      const float score = __fadd_rn(qf, u);
      if (a == 0 || score > best_score) {  // first maximum wins: lowest-index tie-break
        best_score = score;
        best = a;
      }
    }
    int b_child = 0, b_n = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
      if (a == best) {
        b_child = rec.child(a);
        b_n = rec.N(a);
      }
    if (path_ent != nullptr && depth < kPathCap) path_ent[depth] = (uint32_t)e | ((uint32_t)best << 16);
    ++depth;
    if (b_child == (int)HMZ_NO_CHILD) {
      leaf.parent = e;
      leaf.action = best;
      leaf.depth = depth;
      return leaf;
    }
    e = b_child;
    n_parent = b_n;
  }
}

// Backup of a path of depth <= 8 recorded in path_ent: all levels' slots are loaded up front (the
// loads are independent), then the leaf-to-root float64 recurrence runs in registers.
__device__ __forceinline__ void backup_thread8(hmz_node_t* nodes, const uint32_t* __restrict__ path_ent, int depth, int pe,
                                               int pa, int sim, float r, double value, double discount, double& root_w,
                                               double& mn, double& mx) {
  const uint4 p0 = *reinterpret_cast<const uint4*>(path_ent);
  const uint4 p1 = depth > 4 ? *reinterpret_cast<const uint4*>(path_ent + 4) : make_uint4(0, 0, 0, 0);
  const uint32_t ent[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
  double w_sum[8], rwd[8];
  int n[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w_sum[k] = 0.0; rwd[k] = 0.0; n[k] = 0;
    if (k < depth) {
      const hmz_node_t* rec = nodes + (ent[k] & 0xFFFFu);
      const int a = (int)(ent[k] >> 16);
      w_sum[k] = rec->W[a];
      n[k] = rec->N[a];
      rwd[k] = (k == depth - 1) ? (double)r : (double)rec->rwd[a];
    }
  }
  nodes[pe].rwd[pa] = r;
  nodes[pe].child[pa] = (uint16_t)(sim + 1);
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    if (k < depth) {
      hmz_node_t* rec = nodes + (ent[k] & 0xFFFFu);
      const int a = (int)(ent[k] >> 16);
      const double ws = __dadd_rn(w_sum[k], value);
      const int nn = n[k] + 1;
      rec->W[a] = ws;
      rec->N[a] = (uint16_t)nn;
      minmax_update(__dadd_rn(rwd[k], __dmul_rn(discount, __ddiv_rn(ws, (double)nn))), mn, mx);
      value = __dadd_rn(rwd[k], __dmul_rn(discount, value));
    }
  }
  root_w = __dadd_rn(root_w, value);
  minmax_update(__dadd_rn(0.0, __dmul_rn(discount, __ddiv_rn(root_w, (double)(sim + 1)))), mn, mx);
}

__device__ __forceinline__ void write_fresh_record_thread(hmz_node_t* rec, const float* __restrict__ pr, int parent,
                                                          int parent_action) {
  uint4* dst = reinterpret_cast<uint4*>(rec);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  dst[0] = z; dst[1] = z; dst[2] = z;
  dst[3] = make_uint4(__float_as_uint(pr[0]), __float_as_uint(pr[1]), __float_as_uint(pr[2]), __float_as_uint(pr[3]));
  dst[4] = make_uint4(__float_as_uint(pr[4]), __float_as_uint(pr[5]), 0u, 0u);
  dst[5] = z;
  dst[6] = make_uint4(0u, 0u, 0u, 0xFFFFFFFFu);
  dst[7] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, (uint32_t)parent | ((uint32_t)parent_action << 16), 0u);
}

}  // namespace hmz
