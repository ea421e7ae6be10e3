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

// Node.best_child repeated from the root until an unexpanded child (MCTS/mcts.py:80-86,
// MCTS/node.py:72-123).  Called by all 32 lanes of a warp; `lane8` is the lane inside the 8-lane
// segment that owns this search, `valid` is false for padding segments.
//   root_n        root.N  (= number of completed simulations)
//   root_prior64  float64 root priors when the root was Dirichlet-noised, else nullptr
__device__ __forceinline__ Leaf select_leaf(const hmz_node_t* __restrict__ nodes, const double* __restrict__ root_prior64,
                                            double mn, double mx, int root_n, const double* __restrict__ ucb_table,
                                            double discount, int lane8, bool valid, uint8_t* __restrict__ path_out,
                                            int path_cap) {
  const unsigned full = 0xffffffffu;
  const bool normalise = mx > mn;  // MinMaxStats.normalize (MCTS/utils_mcts.py:12-16)
  const double range = __dsub_rn(mx, mn);
  const int seg_base = (threadIdx.x & 31) & ~7;
  int e = 0, n_parent = root_n, depth = 0;
  bool active = valid;
  Leaf leaf{0, 0, 0};
  while (__any_sync(full, active)) {
    float score = -INFINITY;
    int c_n = 0, c_child = (int)HMZ_NO_CHILD;
    if (active && lane8 < 6) {
      const hmz_node_t* rec = nodes + e;
      const double w_sum = rec->W[lane8];
      const float prior = rec->prior[lane8];
      const float rwd = rec->rwd[lane8];
      c_n = rec->N[lane8];
      c_child = rec->child[lane8];
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
      const float os = __shfl_xor_sync(full, score, off);
      const int ob = __shfl_xor_sync(full, best, off);
      if (os > score || (os == score && ob < best)) {
        score = os;
        best = ob;
      }
    }
    const int b_child = __shfl_sync(full, c_child, seg_base | best);
    const int b_n = __shfl_sync(full, c_n, seg_base | best);
    if (active) {
      if (path_out != nullptr && lane8 == 0 && depth < path_cap) path_out[depth] = (uint8_t)best;
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

}  // namespace hmz
