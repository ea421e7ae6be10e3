// Device-side building blocks of the tree store (hmz_tree.cu).
//
// A search is owned by a PAIR of lanes (16 searches per warp).  Lane h of the pair owns half h of
// every 128-byte node record: the three 16-byte child slots 3h..3h+2 and their priors, i.e. four
// 128-bit loads of 64 contiguous bytes, evaluated with 3-way instruction-level parallelism; one
// shuffle merges the two lanes' candidates.  Why pairs: ncu / clock timelines of the earlier forms
// showed (a) 8 lanes per search (one child per lane, shuffle arg-max) issue-bound — ~36 warp
// instructions per search-level, a quarter of the lanes idle; (b) one thread per search bound by
// the L1 tag stage, which serves one 128-byte line per cycle: 8 x LDG.128 with 32 different lines
// each = 256 cycles per warp-level.  Pairs need ~14 instructions and 4 line-cycles per search-level.
//
// Arithmetic contract (bit-exact vs the reference):
//   Q, W, value, min/max  : IEEE float64, every op an explicit __d*_rn so nvcc cannot contract
//                           the reference's separate multiply and add into an FMA;
//   f32(Q) + f32(U)       : float32 add after rounding each term (MCTS/node.py:83,103,123);
//   U = prior * w         : float32 x float32 for a float32 prior (NumPy >= 2 weak scalars),
//                           float64 product then rounded at a Dirichlet-noised root (:122);
//   ties                  : lowest action index (the sanctioned replacement of :86).
#pragma once
#include <math.h>

#include "hmz_common.cuh"

namespace hmz {

static_assert(sizeof(hmz_child_t) == 16 && sizeof(hmz_half_t) == 64 && sizeof(hmz_node_t) == 128,
              "node record layout (include/hmz.h) must be 2 x 64-byte halves of 3 x 16-byte child slots");

constexpr int kPathCap = 32;  // path levels recorded for the backup (deeper paths walk parent links)

struct Leaf {
  int parent;  // record of the leaf's parent
  int action;  // action from that parent to the (unexpanded) leaf
  int depth;   // number of best_child steps taken (>= 1)
};

// 16-byte child slot <-> registers
struct Slot {
  double W;
  float rwd;
  int n, child;
  __device__ __forceinline__ static Slot unpack(const uint4& q) {
    Slot s;
    s.W = __hiloint2double((int)q.y, (int)q.x);
    s.rwd = __uint_as_float(q.z);
    s.n = (int)(q.w & 0xFFFFu);
    s.child = (int)(q.w >> 16);
    return s;
  }
  __device__ __forceinline__ uint4 pack() const {
    return make_uint4((uint32_t)__double2loint(W), (uint32_t)__double2hiint(W), __float_as_uint(rwd),
                      (uint32_t)n | ((uint32_t)child << 16));
  }
};

__device__ __forceinline__ uint4* slot_ptr(hmz_node_t* nodes, int e, int a) {
  return reinterpret_cast<uint4*>(&nodes[e].h[a / 3].c[a % 3]);
}

// Node.expand (MCTS/node.py:30-51) of a fresh node: six children with priors `pr`, N = 0, W = 0,
// rwd = 0, no expanded grandchildren.  Lane `half` of the pair writes its 64-byte half.
__device__ __forceinline__ void write_fresh_half(hmz_node_t* rec, int half, const float* __restrict__ pr, int parent,
                                                 int parent_action) {
  uint4* dst = reinterpret_cast<uint4*>(&rec->h[half]);
  const uint4 empty = make_uint4(0u, 0u, 0u, 0xFFFF0000u);  // W = 0, rwd = 0, N = 0, child = HMZ_NO_CHILD
  dst[0] = empty;
  dst[1] = empty;
  dst[2] = empty;
  dst[3] = make_uint4(__float_as_uint(pr[3 * half]), __float_as_uint(pr[3 * half + 1]), __float_as_uint(pr[3 * half + 2]),
                      half == 0 ? ((uint32_t)parent | ((uint32_t)parent_action << 16)) : 0u);
}

__device__ __forceinline__ void minmax_update(double x, double& mn, double& mx) {
  if (x > mx) mx = x;  // python max(maximum, value): value only when strictly greater
  if (x < mn) mn = x;
}

// Node.best_child repeated from the root until an unexpanded child (MCTS/mcts.py:80-86,
// MCTS/node.py:72-123), for the pair of lanes that owns the search (`half` = lane & 1).
//   root_n        root.N  (= number of completed simulations)
//   root_prior64  float64 root priors when the root was Dirichlet-noised, else nullptr
//   path_out      nullable [path_cap] bytes: chosen action per level (diagnostics)
//   path_ent      nullable [kPathCap] words: (record | action << 16) per level, for the backup
__device__ __forceinline__ Leaf select_leaf(const hmz_node_t* nodes, const double* __restrict__ root_prior64, double mn,
                                            double mx, int root_n, const double* __restrict__ ucb_table,
                                            double discount, int half, uint8_t* __restrict__ path_out, int path_cap,
                                            uint32_t* __restrict__ path_ent) {
  const int lane = threadIdx.x & 31;
  const unsigned pair = 3u << (lane & ~1);
  const bool normalise = mx > mn;  // MinMaxStats.normalize (MCTS/utils_mcts.py:12-16)
  const double range = __dsub_rn(mx, mn);
  int e = 0, n_parent = root_n, depth = 0;
  Leaf leaf{0, 0, 0};
  while (true) {
    const uint4* hp = reinterpret_cast<const uint4*>(&nodes[e].h[half]);
    const uint4 q0 = hp[0], q1 = hp[1], q2 = hp[2], q3 = hp[3];
    const double tn = ucb_table[n_parent];
    float best_score = 0.f;
    int best = 0, best_child = 0, best_n = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const Slot c = Slot::unpack(j == 0 ? q0 : (j == 1 ? q1 : q2));
      const float prior = __uint_as_float(j == 0 ? q3.x : (j == 1 ? q3.y : q3.z));
      float qf = 0.0f;  // child_Q: 0 for unvisited children (node.py:98-102)
      if (c.n > 0) {
        double q = __dadd_rn((double)c.rwd, __dmul_rn(discount, __ddiv_rn(c.W, (double)c.n)));
        if (normalise) q = __ddiv_rn(__dsub_rn(q, mn), range);
        qf = __double2float_rn(q);
      }
      // child_U: w = (log((N+c_base+1)/c_base) + c_init) * sqrt(N) / (child.N + 1)  (node.py:114-121)
      const double w = __ddiv_rn(tn, (double)(c.n + 1));
      float u;
      if (e == 0 && root_prior64 != nullptr)
        u = __double2float_rn(__dmul_rn(root_prior64[3 * half + j], w));  // float64 prior: product in float64
      else
        u = __fmul_rn(prior, __double2float_rn(w));  // float32 prior: weak scalar -> float32 product
      const float score = __fadd_rn(qf, u);          // node.py:83 on float32 arrays
      if (j == 0 || score > best_score) {             // first maximum wins inside the half
        best_score = score;
        best = 3 * half + j;
        best_child = c.child;
        best_n = c.n;
      }
    }
    // merge the two halves: actions 0..2 (lane 0) beat 3..5 (lane 1) on equal scores
    const float other_score = __shfl_xor_sync(pair, best_score, 1);
    const int other_best = __shfl_xor_sync(pair, best, 1);
    const int other_child = __shfl_xor_sync(pair, best_child, 1);
    const int other_n = __shfl_xor_sync(pair, best_n, 1);
    const bool take_other = half == 0 ? (other_score > best_score) : !(best_score > other_score);
    if (take_other) {
      best = other_best;
      best_child = other_child;
      best_n = other_n;
    }
    if (half == 0) {
      if (path_out != nullptr && depth < path_cap) path_out[depth] = (uint8_t)best;
      if (path_ent != nullptr && depth < kPathCap) path_ent[depth] = (uint32_t)e | ((uint32_t)best << 16);
    }
    ++depth;
    if (best_child == (int)HMZ_NO_CHILD) {
      leaf.parent = e;
      leaf.action = best;
      leaf.depth = depth;
      return leaf;
    }
    e = best_child;
    n_parent = best_n;
  }
}

// node.expand bookkeeping on the parent slot + Node.backup (MCTS/node.py:53-70) from the leaf to
// the root by walking parent links (any depth).  `value` enters as the network value of the new node.
__device__ __forceinline__ void backup_walk(hmz_node_t* nodes, int pe, int pa, int sim, float r, double value,
                                            double discount, double& root_w, double& mn, double& mx) {
  int e = pe, a = pa;
  bool leaf = true;
  while (true) {
    uint4* sp = slot_ptr(nodes, e, a);
    Slot c = Slot::unpack(*sp);
    if (leaf) {  // leaf.rwd = reward, the leaf is now the expanded record sim + 1 (node.py:44-49)
      c.rwd = r;
      c.child = sim + 1;
      leaf = false;
    }
    c.W = __dadd_rn(c.W, value);  // current.W += value
    c.n += 1;                     // current.N += 1
    *sp = c.pack();
    const double rwd = (double)c.rwd;
    minmax_update(__dadd_rn(rwd, __dmul_rn(discount, __ddiv_rn(c.W, (double)c.n))), mn, mx);
    value = __dadd_rn(rwd, __dmul_rn(discount, value));  // value = rwd + discount * value
    if (e == 0) break;
    a = nodes[e].h[0].parent_action;
    e = nodes[e].h[0].parent;
  }
  // the root itself: rwd = 0.0 (MCTS/mcts.py:69), N = sim + 1 after this backup
  root_w = __dadd_rn(root_w, value);
  minmax_update(__dadd_rn(0.0, __dmul_rn(discount, __ddiv_rn(root_w, (double)(sim + 1)))), mn, mx);
}

// Same backup for a path of depth <= 8 recorded by select_leaf: the slots are loaded four levels at a
// time up front (independent 16-byte loads: one memory round trip per batch instead of one per level),
// then the leaf-to-root float64 recurrence runs in registers.  Operation for operation identical to
// backup_walk, so results are bit-identical.
__device__ __forceinline__ void backup_batch4(hmz_node_t* nodes, const uint4& ent4, int k0, int depth, int sim, float r,
                                              double& value, double discount, double& mn, double& mx) {
  const uint32_t ent[4] = {ent4.x, ent4.y, ent4.z, ent4.w};
  uint4 raw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (k0 + j < depth) raw[j] = *slot_ptr(nodes, (int)(ent[j] & 0xFFFFu), (int)(ent[j] >> 16));
#pragma unroll
  for (int j = 3; j >= 0; --j) {
    if (k0 + j < depth) {
      Slot c = Slot::unpack(raw[j]);
      if (k0 + j == depth - 1) {  // the leaf slot: Node.expand bookkeeping on the parent (node.py:44-49)
        c.rwd = r;
        c.child = sim + 1;
      }
      c.W = __dadd_rn(c.W, value);  // current.W += value
      c.n += 1;                     // current.N += 1
      *slot_ptr(nodes, (int)(ent[j] & 0xFFFFu), (int)(ent[j] >> 16)) = c.pack();
      const double rwd = (double)c.rwd;
      minmax_update(__dadd_rn(rwd, __dmul_rn(discount, __ddiv_rn(c.W, (double)c.n))), mn, mx);
      value = __dadd_rn(rwd, __dmul_rn(discount, value));  // value = rwd + discount * value
    }
  }
}

__device__ __forceinline__ void backup_path8(hmz_node_t* nodes, const uint32_t* __restrict__ path_ent, int depth, int sim,
                                             float r, double value, double discount, double& root_w, double& mn,
                                             double& mx) {
  if (depth > 4) backup_batch4(nodes, *reinterpret_cast<const uint4*>(path_ent + 4), 4, depth, sim, r, value, discount, mn, mx);
  backup_batch4(nodes, *reinterpret_cast<const uint4*>(path_ent), 0, depth, sim, r, value, discount, mn, mx);
  root_w = __dadd_rn(root_w, value);
  minmax_update(__dadd_rn(0.0, __dmul_rn(discount, __ddiv_rn(root_w, (double)(sim + 1)))), mn, mx);
}

}  // namespace hmz
