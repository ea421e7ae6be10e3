// Device-side building blocks of the Tower-of-Hanoi env kernels (hmz_env.cu) — shared with the fused self-play
// move kernel (hmz_selfplay.cu).  Restates env/hanoi.py:47-84 (step), :123-139 (_move_allowed), :141-151
// (_get_moved_state) of the reference on packed 32-bit env words (layout in include/hmz.h).
#pragma once
#include "hmz_common.cuh"

namespace hmz {

struct EnvCfg {
  uint32_t even_mask;   // 0x55555555 restricted to the 2N state bits
  uint32_t state_mask;  // (1 << 2N) - 1
  uint32_t goal_word;   // goal_peg replicated on every disk
  uint32_t reset_word;
  uint32_t max_steps;
  int shift;            // 2N: where the step counter starts
  int auto_reset;
};

// The kernels are issue-bound, not memory-bound (ncu, round 2: 70 % issue-active at 0.65 of the HBM roofline with the
// first, branchy form: ~85 SASS instructions per env, 13 divergent regions), so everything below is straight-line
// select / logic code with no find-first-set and no data-dependent branch.
//
// "Key" of a peg: (lowest set bit of the mask of disks on the peg) - 1, i.e. 2^(2 * top disk) - 1, and 0xFFFFFFFF for an
// empty peg: keys order the pegs exactly as their top disks do, with "empty" = +inf.
__device__ __forceinline__ void peg_keys(uint32_t st, uint32_t even_mask, uint32_t& k0, uint32_t& k1, uint32_t& k2) {
  const uint32_t lo = st & even_mask, hi = (st >> 1) & even_mask;
  const uint32_t m0 = even_mask & ~(lo | hi), m1 = lo & ~hi, m2 = hi & ~lo;
  k0 = (m0 & (0u - m0)) - 1u;
  k1 = (m1 & (0u - m1)) - 1u;
  k2 = (m2 & (0u - m2)) - 1u;
}

// bit a = move a allowed; actions 0:(0,1) 1:(0,2) 2:(1,0) 3:(1,2) 4:(2,0) 5:(2,1)  (env/hanoi.py:39-41).
// A move f->t is allowed iff peg f is non-empty and (peg t is empty or its top disk is larger),
// i.e. key(f) < key(t) (env/hanoi.py:123-139).
__device__ __forceinline__ uint32_t legal_bits(uint32_t k0, uint32_t k1, uint32_t k2) {
  return (uint32_t)(k0 < k1) | ((uint32_t)(k0 < k2) << 1) | ((uint32_t)(k1 < k0) << 2) | ((uint32_t)(k1 < k2) << 3) |
         ((uint32_t)(k2 < k0) << 4) | ((uint32_t)(k2 < k1) << 5);
}

struct StepOut {
  uint32_t word;      // new env word (state | counter << shift), after optional auto-reset
  uint32_t obs_word;  // state the returned observation encodes
  float reward;
  uint32_t flags;
};

// The tail of TowersOfHanoi.step once the move's legality and the moved disk's bit are known (env/hanoi.py:56-80):
//   legal, next != goal -> (next, 0, not done);  legal, next == goal -> (obs = goal, 100, done, STORED STATE UNCHANGED,
//   counter 0);  illegal -> (same, -0.1, not done);  counter == max_steps -> done, counter 0 (can co-occur with illegal).
__device__ __forceinline__ StepOut step_finish(uint32_t word, uint32_t st, bool legal, uint32_t moved, const EnvCfg& c) {
  uint32_t ctr = (word >> c.shift) + 1u;  // :56 — counted for illegal moves too
  const bool goal = legal & (moved == c.goal_word);
  const uint32_t stored = (legal & !goal) ? moved : st;  // :65-69 — the goal state is returned but not stored
  StepOut o;
  o.obs_word = legal ? moved : st;
  o.reward = goal ? 100.0f : (legal ? 0.0f : -0.1f);  // -0.1f: float32 image of the python double -100/1000 (:72)
  uint32_t flags = (goal ? (HMZ_FLAG_DONE | HMZ_FLAG_GOAL) : 0u) | (legal ? 0u : HMZ_FLAG_ILLEGAL);
  ctr = goal ? 0u : ctr;
  const bool trunc = ctr == c.max_steps;  // :77-80
  flags |= trunc ? (HMZ_FLAG_DONE | HMZ_FLAG_TRUNC) : 0u;
  ctr = trunc ? 0u : ctr;
  uint32_t w = stored | (ctr << c.shift);
  if (c.auto_reset) w = (flags & HMZ_FLAG_DONE) ? c.reset_word : w;
  o.word = w;
  o.flags = flags;
  return o;
}

// TowersOfHanoi.step for a GIVEN action: only the source and the target peg matter.  XOR-ing the state with the peg
// number replicated over all disks turns "disk on that peg" into a zero digit, so one mask expression serves any peg.
__device__ __forceinline__ StepOut step_word(uint32_t word, uint32_t action, const EnvCfg& c) {
  const uint32_t st = word & c.state_mask;
  const uint32_t a = action > 5u ? 5u : action;
  const uint32_t f = a >> 1, t = (0x489u >> (2u * a)) & 3u;  // moves[a] = (f, t) (env/hanoi.py:39-41)
  const uint32_t xf = st ^ (f * c.even_mask), xt = st ^ (t * c.even_mask);
  const uint32_t mf = ~(xf | (xf >> 1)) & c.even_mask, mt = ~(xt | (xt >> 1)) & c.even_mask;
  const uint32_t lf = mf & (0u - mf);  // bit of the top disk of the source peg (0: the peg is empty)
  // _move_allowed (:123-139): source non-empty and no disk of the target peg below its top disk
  const bool legal = (action <= 5u) & (lf != 0u) & ((mt & (lf - 1u)) == 0u);
  const uint32_t moved = st ^ ((f ^ t) * lf);  // _get_moved_state (:141-151): one digit changes from f to t
  return step_finish(word, st, legal, moved, c);
}

// One step on a uniformly random LEGAL move.  Tower of Hanoi has exactly two or three legal moves in any state: the
// smallest disk (disk 0, on peg p0) can always go to either of the other two pegs q1, q2; between those two pegs the
// smaller of the two top disks can move onto the other (no third move when both are empty, i.e. when every disk sits
// on p0).  So the legal set is known from disk 0's peg and one comparison — no mask over all six moves, no k-th-set-bit
// search: the random-move kernels are integer-bound (ncu: the HBM roofline is not what limits them).
// Index in the move table [(0,1),(0,2),(1,0),(1,2),(2,0),(2,1)] (env/hanoi.py:39-41): a = 2 f + (t > f ? t - 1 : t).
__device__ __forceinline__ StepOut step_random_word(uint32_t word, uint32_t rnd, const EnvCfg& c, uint32_t& action) {
  const uint32_t st = word & c.state_mask;
  const uint32_t p0 = st & 3u;
  const uint32_t q1 = p0 == 2u ? 0u : p0 + 1u, q2 = p0 == 0u ? 2u : p0 - 1u;
  const uint32_t x1 = st ^ (q1 * c.even_mask), x2 = st ^ (q2 * c.even_mask);
  const uint32_t m1 = ~(x1 | (x1 >> 1)) & c.even_mask, m2 = ~(x2 | (x2 >> 1)) & c.even_mask;
  const uint32_t l1 = m1 & (0u - m1), l2 = m2 & (0u - m2);  // bits of the top disks of q1 / q2 (0: empty peg)
  const uint32_t k1 = l1 - 1u, k2 = l2 - 1u;                // keys: empty = 0xFFFFFFFF
  const uint32_t n_legal = 2u + (uint32_t)(k1 != k2);
  const uint32_t idx = __umulhi(rnd, n_legal);              // uniform over the legal moves
  const bool first = k1 < k2;                               // the third move goes q1 -> q2, else q2 -> q1
  const uint32_t f = idx < 2u ? p0 : (first ? q1 : q2);
  const uint32_t t = idx == 0u ? q1 : (idx == 1u ? q2 : (first ? q2 : q1));
  const uint32_t lf = idx < 2u ? 1u : (first ? l1 : l2);    // bit of the disk that moves (disk 0: bit 0)
  action = 2u * f + t - (uint32_t)(t > f);
  return step_finish(word, st, true, st ^ ((f ^ t) * lf), c);  // the chosen move is legal by construction
}

// Host: fills an EnvCfg after validating the shape (HMZ_ERR_* with a message on failure).
int make_env_cfg(EnvCfg& c, int n_disks, int max_steps, int goal_peg, int auto_reset, uint32_t reset_word);

}  // namespace hmz
