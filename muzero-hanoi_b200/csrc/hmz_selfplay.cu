// Self-play glue kernels: on-device Dirichlet / uniform draws (throughput mode) and the
// per-move trajectory record (Muzero._play_game, Muzero.py:153-207 of the reference).
#include "hmz_common.cuh"

namespace hmz {

struct PhiloxStream {  // sequential 32-bit draws from one (key, item) Philox stream
  uint32_t c0, c1, c2, c3, k0, k1;
  Philox4 buf;
  int have;
  __device__ PhiloxStream(uint64_t seed, uint64_t item, uint64_t counter)
      : c0((uint32_t)item), c1((uint32_t)(item >> 32)), c2((uint32_t)counter), c3((uint32_t)(counter >> 32) << 8),
        k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), have(0) {}
  __device__ uint32_t next() {
    if (have == 0) {
      buf = philox4x32_10(c0, c1, c2, c3, k0, k1);
      ++c3;  // low 8 bits of c3: up to 256 blocks per (item, counter)
      have = 4;
    }
    --have;
    return have == 3 ? buf.x : (have == 2 ? buf.y : (have == 1 ? buf.z : buf.w));
  }
  __device__ double uniform() {  // (0, 1): 53 random bits, never exactly 0
    const uint64_t hi = next(), lo = next();
    return ((double)(((hi << 32) | lo) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ double normal() {  // Box-Muller
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
  // Marsaglia & Tsang (2000); for alpha < 1: gamma(alpha) = gamma(alpha + 1) * U^(1/alpha)
  __device__ double gamma(double alpha) {
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double out;
    for (;;) {
      double x = normal(), t = 1.0 + c * x;
      if (t <= 0.0) continue;
      t = t * t * t;
      const double u = uniform();
      if (log(u) < 0.5 * x * x + d - d * t + d * log(t)) {
        out = d * t;
        break;
      }
    }
    if (alpha < 1.0) out *= pow(uniform(), 1.0 / alpha);
    return out;
  }
};

__global__ void __launch_bounds__(256) rng_dirichlet(double* __restrict__ out, int64_t n, double alpha, uint64_t seed,
                                                    uint64_t counter) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    PhiloxStream rng(seed ^ 0x4449524943484C45ull, (uint64_t)i, counter);
    double g[6], sum = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      g[a] = rng.gamma(alpha);
      sum += g[a];
    }
    if (!(sum > 0.0)) {  // all six gammas underflowed (alpha tiny): fall back to one-hot on a random action
      const int pick = (int)(rng.uniform() * 6.0);
#pragma unroll
      for (int a = 0; a < 6; ++a) g[a] = (a == pick) ? 1.0 : 0.0;
      sum = 1.0;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) out[i * 6 + a] = g[a] / sum;
  }
}

__global__ void __launch_bounds__(256) rng_uniform(double* __restrict__ out, int64_t n, uint64_t seed,
                                                  uint64_t counter) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)counter, (uint32_t)(counter >> 32),
                              (uint32_t)seed ^ 0x554E4946u, (uint32_t)(seed >> 32));
    out[i] = (double)((((uint64_t)r.x << 32) | r.y) >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
  }
}

__global__ void __launch_bounds__(256) traj_record(const uint32_t* __restrict__ words, const int32_t* __restrict__ action,
                                                  const int32_t* __restrict__ visits, const double* __restrict__ root_q,
                                                  uint32_t* __restrict__ t_state, uint8_t* __restrict__ t_action,
                                                  uint16_t* __restrict__ t_visits, float* __restrict__ t_root_q,
                                                  uint8_t* __restrict__ action_u8, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t a = (uint8_t)action[i];
    if (t_state) t_state[i] = words[i];
    if (t_action) t_action[i] = a;
    if (action_u8) action_u8[i] = a;
    if (t_root_q) t_root_q[i] = (float)root_q[i];
    if (t_visits) {
#pragma unroll
      for (int k = 0; k < 6; ++k) t_visits[i * 6 + k] = (uint16_t)visits[i * 6 + k];
    }
  }
}

}  // namespace hmz

using namespace hmz;

extern "C" {

int hmz_rng_dirichlet(double* out, int64_t n, double alpha, uint64_t seed, uint64_t counter, void* stream) {
  ProfScope prof_scope(HMZ_PROF_OTHER, stream);
  if (n == 0) return HMZ_OK;
  if (!out || n < 0 || !(alpha > 0.0)) return fail(HMZ_ERR_INVALID, "hmz_rng_dirichlet: bad arguments");
  rng_dirichlet<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(out, n, alpha, seed, counter);
  return check_launch("rng_dirichlet");
}

int hmz_rng_uniform(double* out, int64_t n, uint64_t seed, uint64_t counter, void* stream) {
  ProfScope prof_scope(HMZ_PROF_OTHER, stream);
  if (n == 0) return HMZ_OK;
  if (!out || n < 0) return fail(HMZ_ERR_INVALID, "hmz_rng_uniform: bad arguments");
  rng_uniform<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(out, n, seed, counter);
  return check_launch("rng_uniform");
}

int hmz_traj_record(const uint32_t* words, const int32_t* action, const int32_t* visits, const double* root_q,
                    uint32_t* traj_state, uint8_t* traj_action, uint16_t* traj_visits, float* traj_root_q,
                    uint8_t* action_u8_out, int64_t n, void* stream) {
  ProfScope prof_scope(HMZ_PROF_OTHER, stream);
  if (n == 0) return HMZ_OK;
  if (!action || n < 0 || (traj_state && !words) || (traj_visits && !visits) || (traj_root_q && !root_q))
    return fail(HMZ_ERR_INVALID, "hmz_traj_record: bad arguments");
  traj_record<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(words, action, visits, root_q, traj_state,
                                                                      traj_action, traj_visits, traj_root_q,
                                                                      action_u8_out, n);
  return check_launch("traj_record");
}

// One move of every game (Muzero._play_game loop body, Muzero.py:165-186): root inference -> Dirichlet mix ->
// n_simulations fused simulations -> root policy + sampled action -> trajectory / episode record -> env step.
// A composition of the entry points above: no host synchronisation, everything on `stream`.
int hmz_selfplay_move(const hmz_selfplay_t* sp, uint64_t move_index, void* stream) {
  if (!sp) return fail(HMZ_ERR_INVALID, "hmz_selfplay_move: null descriptor");
  const hmz_search_t* s = &sp->search;
  const int64_t B = s->n_searches;
  if (B == 0) return HMZ_OK;
  if (!sp->weights || !sp->words || !sp->p0 || !sp->v0 || !sp->uniform || !sp->visits || !sp->root_q || !sp->action ||
      !sp->action_u8 || !sp->step_reward || !sp->step_flags || !sp->ucb_table || s->n_searches < 0)
    return fail(HMZ_ERR_INVALID, "hmz_selfplay_move: null buffer in descriptor");
  const bool use_noise = sp->dirichlet_alpha > 0.0 && sp->exploration_eps > 0.0;
  if (use_noise && !sp->noise) return fail(HMZ_ERR_INVALID, "hmz_selfplay_move: noise buffer required when dirichlet_alpha > 0");
  if (s->root_prior_is_f64 != (use_noise ? 1 : 0))
    return fail(HMZ_ERR_INVALID, "hmz_selfplay_move: search.root_prior_is_f64 must be %d", use_noise ? 1 : 0);
  if (int rc = hmz_net_initial(sp->weights, sp->mode, sp->n_disks, sp->words, nullptr, s->latents, s->n_records, s->latent_dtype,
                               sp->p0, sp->v0, B, stream))
    return rc;
  if (use_noise)
    if (int rc = hmz_rng_dirichlet(sp->noise, B, sp->dirichlet_alpha, sp->seed, move_index, stream)) return rc;
  if (int rc = hmz_rng_uniform(sp->uniform, B, sp->seed, move_index, stream)) return rc;
  if (int rc = hmz_search_begin_p0(s, sp->p0, use_noise ? sp->noise : nullptr, sp->exploration_eps, stream)) return rc;
  if (int rc = hmz_search_run(s, sp->weights, sp->mode, sp->n_simulations, sp->ucb_table, sp->discount, stream)) return rc;
  if (int rc = hmz_search_root_policy(s, sp->n_simulations, sp->temperature, 0, sp->uniform, nullptr, sp->visits, nullptr, sp->root_q,
                                      sp->action, stream))
    return rc;
  if (int rc = hmz_traj_record(sp->words, sp->action, sp->visits, sp->root_q, sp->traj_state, sp->traj_action, sp->traj_visits,
                               sp->traj_root_q, sp->action_u8, B, stream))
    return rc;
  if (sp->ep_state) {
    if (int rc = hmz_episode_record(sp->words, sp->action, sp->visits, sp->root_q, sp->n_disks, sp->ep_t_max, B, sp->ep_state,
                                    sp->ep_action, sp->ep_visits, sp->ep_root_q, sp->ep_cur_slot, nullptr, stream))
      return rc;
  }
  if (int rc = hmz_env_step(sp->words, sp->action_u8, sp->step_reward, sp->step_flags, nullptr, B, sp->n_disks, sp->max_steps,
                            sp->goal_peg, 1, sp->reset_word, stream))
    return rc;
  if (sp->ep_state)
    if (int rc = hmz_episode_close(sp->step_flags, sp->ep_cur_slot, B, sp->ep_flags, sp->ep_len, stream)) return rc;
  return HMZ_OK;
}

}  // extern "C"
