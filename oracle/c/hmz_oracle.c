/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's env step and MCTS tree
 * arithmetic, used as the fast checker for the large GPU parity cases (4,096+ searches).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may load this library.
 *
 * Follows (file:line under the reference repository):
 *   env/hanoi.py:47-84 (step), :123-139 (_move_allowed), :141-151 (_get_moved_state)
 *   env/hanoi_utils.py:4-26 (hanoi_solver)
 *   MCTS/mcts.py:71-109 (simulation loop), MCTS/node.py:53-123 (backup, best_child, child_Q,
 *   child_U), MCTS/utils_mcts.py:8-16 (MinMaxStats)
 * Pinned against the golden fixtures generated from the unmodified reference by
 * tests/test_oracle_golden.py::test_c_oracle_*.  Compile with -ffp-contract=off: the reference
 * rounds every float64 multiply and add separately.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int MOVE_FROM[6] = {0, 0, 1, 1, 2, 2};
static const int MOVE_TO[6] = {1, 2, 0, 2, 0, 1};

static int top_disk(uint32_t st, int n, int peg) {
  for (int d = 0; d < n; ++d)
    if (((st >> (2 * d)) & 3u) == (uint32_t)peg) return d;
  return -1;
}

static int move_allowed(uint32_t st, int n, int a) {
  int tf = top_disk(st, n, MOVE_FROM[a]);
  if (tf < 0) return 0;
  int tt = top_disk(st, n, MOVE_TO[a]);
  return tt < 0 ? 1 : tt > tf;
}

/* Same word layout and flag bits as include/hmz.h. */
void oracle_env_step(uint32_t* words, const uint8_t* actions, float* rewards, uint8_t* flags, uint32_t* obs_words,
                     int64_t n_envs, int n, int max_steps, int goal_peg, int auto_reset, uint32_t reset_word) {
  const int shift = 2 * n;
  const uint32_t mask = (1u << shift) - 1u;
  uint32_t goal = 0;
  for (int d = 0; d < n; ++d) goal |= (uint32_t)goal_peg << (2 * d);
  for (int64_t i = 0; i < n_envs; ++i) {
    uint32_t st = words[i] & mask, ctr = (words[i] >> shift) + 1u, stored = st, obs = st, fl = 0;
    float rw = 0.0f;
    int a = actions[i];
    if (a < 6 && move_allowed(st, n, a)) {
      int d = top_disk(st, n, MOVE_FROM[a]);
      uint32_t moved = (st & ~(3u << (2 * d))) | ((uint32_t)MOVE_TO[a] << (2 * d));
      obs = moved;
      if (moved == goal) { rw = 100.0f; fl = 1u | 4u; ctr = 0; } else stored = moved;
    } else { rw = -0.1f; fl = 2u; }
    if (ctr == (uint32_t)max_steps) { fl |= 1u | 8u; ctr = 0; }
    uint32_t w = stored | (ctr << shift);
    if (auto_reset && (fl & 1u)) w = reset_word;
    words[i] = w; rewards[i] = rw; flags[i] = (uint8_t)fl;
    if (obs_words) obs_words[i] = obs;
  }
}

void oracle_legal_mask(const uint32_t* words, uint8_t* mask, int64_t n_envs, int n) {
  for (int64_t i = 0; i < n_envs; ++i) {
    uint32_t st = words[i] & ((1u << (2 * n)) - 1u);
    uint8_t m = 0;
    for (int a = 0; a < 6; ++a) m |= (uint8_t)(move_allowed(st, n, a) << a);
    mask[i] = m;
  }
}

void oracle_solver(const uint32_t* words, uint32_t* dist, int64_t n_envs, int n, int goal_peg) {
  for (int64_t i = 0; i < n_envs; ++i) {
    uint32_t moves = 0, target = (uint32_t)goal_peg;
    for (int d = n - 1; d >= 0; --d) {
      uint32_t peg = (words[i] >> (2 * d)) & 3u;
      if (peg != target) { moves += 1u << d; target = 3u - target - peg; }
    }
    dist[i] = moves;
  }
}

/* One search with the network outputs of every simulation injected.
 *   prior[6] float64 (exact float32 values unless prior_is_f64), minmax[2] in/out,
 *   r[S], p[S][6], v[S] float32, table[n] = (log((n+19653)/19652)+1.25)*sqrt(n) for n <= S.
 * Outputs visits[6], *root_q, and (optional) leaf_depth[S]. */
static void search_one(int S, double discount, const double* prior, int prior_is_f64, double* minmax, const float* r,
                       const float* p, const float* v, const double* table, int32_t* visits, double* root_q,
                       uint16_t* leaf_depth) {
  const int E = S + 1;
  int32_t* cN = calloc((size_t)E * 6, sizeof(int32_t));
  double* cW = calloc((size_t)E * 6, sizeof(double));
  double* cR = calloc((size_t)E * 6, sizeof(double));
  double* cP = calloc((size_t)E * 6, sizeof(double));
  int32_t* cE = malloc((size_t)E * 6 * sizeof(int32_t));
  int32_t* pe_of = malloc((size_t)E * sizeof(int32_t));
  int32_t* pa_of = malloc((size_t)E * sizeof(int32_t));
  for (int i = 0; i < E * 6; ++i) cE[i] = -1;
  for (int a = 0; a < 6; ++a) cP[a] = prior[a];
  double mn = minmax[0], mx = minmax[1], root_w = 0.0;
  for (int s = 0; s < S; ++s) {
    int e = 0, n_parent = s, depth = 0, best = 0;
    for (;;) {
      float best_score = 0.0f;
      for (int a = 0; a < 6; ++a) {
        int n = cN[e * 6 + a];
        float qf = 0.0f;
        if (n > 0) {
          double q = cR[e * 6 + a] + discount * (cW[e * 6 + a] / (double)n);
          if (mx > mn) q = (q - mn) / (mx - mn);
          qf = (float)q;
        }
        double w = table[n_parent] / (double)(n + 1);
        float u = (e == 0 && prior_is_f64) ? (float)(cP[e * 6 + a] * w) : (float)cP[e * 6 + a] * (float)w;
        float score = qf + u;
        if (a == 0 || score > best_score) { best = a; best_score = score; }
      }
      ++depth;
      if (cE[e * 6 + best] < 0) break;
      n_parent = cN[e * 6 + best];
      e = cE[e * 6 + best];
    }
    if (leaf_depth) leaf_depth[s] = (uint16_t)depth;
    const int nw = s + 1;
    cE[e * 6 + best] = nw; pe_of[nw] = e; pa_of[nw] = best;
    cR[e * 6 + best] = (double)r[s];
    for (int a = 0; a < 6; ++a) cP[nw * 6 + a] = (double)p[s * 6 + a];
    double value = (double)v[s];
    int ce = e, ca = best;
    for (;;) {
      cW[ce * 6 + ca] += value;
      cN[ce * 6 + ca] += 1;
      double x = cR[ce * 6 + ca] + discount * (cW[ce * 6 + ca] / (double)cN[ce * 6 + ca]);
      if (x > mx) mx = x;
      if (x < mn) mn = x;
      value = cR[ce * 6 + ca] + discount * value;
      if (ce == 0) break;
      int t = ce; ca = pa_of[t]; ce = pe_of[t];
    }
    root_w += value;
    double x = 0.0 + discount * (root_w / (double)(s + 1));
    if (x > mx) mx = x;
    if (x < mn) mn = x;
  }
  for (int a = 0; a < 6; ++a) visits[a] = cN[a];
  *root_q = S > 0 ? root_w / (double)S : 0.0;
  minmax[0] = mn; minmax[1] = mx;
  free(cN); free(cW); free(cR); free(cP); free(cE); free(pe_of); free(pa_of);
}

/* Batch layout matches the GPU engine: r, v [S][B]; p [S][B][6]; prior [B][6]; minmax [B][2]. */
void oracle_search_injected(int64_t B, int S, double discount, const double* prior, int prior_is_f64, double* minmax,
                            const float* r, const float* p, const float* v, const double* table, int32_t* visits,
                            double* root_q, uint16_t* leaf_depth /* [S][B] or NULL */) {
  float* rb = malloc((size_t)S * sizeof(float));
  float* vb = malloc((size_t)S * sizeof(float));
  float* pb = malloc((size_t)S * 6 * sizeof(float));
  uint16_t* db = malloc((size_t)S * sizeof(uint16_t));
  for (int64_t b = 0; b < B; ++b) {
    for (int s = 0; s < S; ++s) {
      rb[s] = r[(int64_t)s * B + b];
      vb[s] = v[(int64_t)s * B + b];
      memcpy(pb + s * 6, p + ((int64_t)s * B + b) * 6, 6 * sizeof(float));
    }
    search_one(S, discount, prior + b * 6, prior_is_f64, minmax + b * 2, rb, pb, vb, table, visits + b * 6, root_q + b,
               db);
    if (leaf_depth)
      for (int s = 0; s < S; ++s) leaf_depth[(int64_t)s * B + b] = db[s];
  }
  free(rb); free(vb); free(pb); free(db);
}
