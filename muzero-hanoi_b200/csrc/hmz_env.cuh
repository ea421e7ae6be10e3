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

// Bit index (2*disk) of the top (smallest) disk on each peg, 0xFFFFFFFF if the peg is empty.
__device__ __forceinline__ void peg_tops(uint32_t st, uint32_t even_mask, uint32_t& t0, uint32_t& t1, uint32_t& t2) {
  uint32_t lo = st & even_mask, hi = (st >> 1) & even_mask;
  t0 = (uint32_t)(__ffs((int)(even_mask & ~(lo | hi))) - 1);
  t1 = (uint32_t)(__ffs((int)(lo & ~hi)) - 1);
  t2 = (uint32_t)(__ffs((int)(hi & ~lo)) - 1);
}

// bit a = move a allowed; actions 0:(0,1) 1:(0,2) 2:(1,0) 3:(1,2) 4:(2,0) 5:(2,1)  (env/hanoi.py:39-41).
// A move f->t is allowed iff peg f is non-empty and (peg t is empty or its top disk is larger),
// i.e. top(f) < top(t) with "empty" = +inf (env/hanoi.py:123-139).
__device__ __forceinline__ uint32_t legal_bits(uint32_t t0, uint32_t t1, uint32_t t2) {
  return (uint32_t)(t0 < t1) | ((uint32_t)(t0 < t2) << 1) | ((uint32_t)(t1 < t0) << 2) | ((uint32_t)(t1 < t2) << 3) |
         ((uint32_t)(t2 < t0) << 4) | ((uint32_t)(t2 < t1) << 5);
}

struct StepOut {
  uint32_t word;      // new env word (state | counter << shift), after optional auto-reset
  uint32_t obs_word;  // state the returned observation encodes
  float reward;
  uint32_t flags;
};

// TowersOfHanoi.step with the peg tops of the state already known (the random-move kernels need them to pick the action).
__device__ __forceinline__ StepOut step_word_tops(uint32_t word, uint32_t action, uint32_t t0, uint32_t t1, uint32_t t2,
                                                  const EnvCfg& c) {
  uint32_t st = word & c.state_mask;
  uint32_t ctr = (word >> c.shift) + 1u;  // env/hanoi.py:56 — counted for illegal moves too
  uint32_t a = action > 5u ? 5u : action;
  uint32_t legal = (action <= 5u) ? ((legal_bits(t0, t1, t2) >> a) & 1u) : 0u;
  uint32_t f = a >> 1;
  uint32_t t = (0x489u >> (2u * a)) & 3u;
  uint32_t tf = f == 0u ? t0 : (f == 1u ? t1 : t2);
  StepOut o;
  o.flags = 0u;
  uint32_t stored = st;
  o.obs_word = st;
  o.reward = 0.0f;
  if (legal) {
    uint32_t moved = st ^ ((f ^ t) << tf);  // env/hanoi.py:141-151: one digit changes
    o.obs_word = moved;
    if (moved == c.goal_word) {  // :65-69 — stored state is NOT updated, counter cleared
      o.reward = 100.0f;
      o.flags = HMZ_FLAG_DONE | HMZ_FLAG_GOAL;
      ctr = 0u;
    } else {
      stored = moved;
    }
  } else {
    o.reward = -0.1f;  // float32 image of the python double -100/1000 (:72)
    o.flags = HMZ_FLAG_ILLEGAL;
  }
  if (ctr == c.max_steps) {  // :77-80
    o.flags |= HMZ_FLAG_DONE | HMZ_FLAG_TRUNC;
    ctr = 0u;
  }
  o.word = stored | (ctr << c.shift);
  if (c.auto_reset && (o.flags & HMZ_FLAG_DONE)) o.word = c.reset_word;
  return o;
}

__device__ __forceinline__ StepOut step_word(uint32_t word, uint32_t action, const EnvCfg& c) {
  uint32_t t0, t1, t2;
  peg_tops(word & c.state_mask, c.even_mask, t0, t1, t2);
  return step_word_tops(word, action, t0, t1, t2, c);
}

// k-th (0-based) set bit of a mask with 2 or 3 bits set.
__device__ __forceinline__ uint32_t kth_set_bit(uint32_t m, uint32_t k) {
  uint32_t m1 = m & (m - 1u);
  uint32_t m2 = m1 & (m1 - 1u);
  uint32_t sel = k == 0u ? m : (k == 1u ? m1 : m2);
  return (uint32_t)(__ffs((int)sel) - 1);
}

__device__ __forceinline__ uint32_t random_legal_action(uint32_t st, uint32_t rnd, const EnvCfg& c) {
  uint32_t t0, t1, t2;
  peg_tops(st, c.even_mask, t0, t1, t2);
  uint32_t m = legal_bits(t0, t1, t2);
  return kth_set_bit(m, __umulhi(rnd, (uint32_t)__popc(m)));
}

// One step on a uniformly random LEGAL move (the peg tops are computed once for the choice and the step).
__device__ __forceinline__ StepOut step_random_word(uint32_t word, uint32_t rnd, const EnvCfg& c, uint32_t& action) {
  uint32_t t0, t1, t2;
  peg_tops(word & c.state_mask, c.even_mask, t0, t1, t2);
  const uint32_t m = legal_bits(t0, t1, t2);
  action = kth_set_bit(m, __umulhi(rnd, (uint32_t)__popc(m)));
  return step_word_tops(word, action, t0, t1, t2, c);
}

// Host: fills an EnvCfg after validating the shape (HMZ_ERR_* with a message on failure).
int make_env_cfg(EnvCfg& c, int n_disks, int max_steps, int goal_peg, int auto_reset, uint32_t reset_word);

}  // namespace hmz
