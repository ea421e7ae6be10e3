"""Drop-in for the reference's MCTS/mcts.py: same constructor, attributes and methods
(cites are reference MCTS/mcts.py:line), as a B=1 view over engine.BatchedMCTS."""
import numpy as np
import torch

from .. import _lib
from ..engine import BatchedMCTS
from .utils_mcts import MinMaxStats


class MCTS:
    def __init__(self, discount, root_dirichlet_alpha, n_simulations, batch_s, device, h_dim=64, clip_grad=True,
                 root_exploration_eps=0.25, known_bounds=[]):
        self.min_max_stats = MinMaxStats()  # :23 — lives as long as this object, never reset
        self.pb_c_base = 19652
        self.pb_c_init = 1.25
        self.discount = discount
        self.root_dirichlet_alpha = root_dirichlet_alpha
        self.root_exploration_eps = root_exploration_eps
        self.n_simulations = n_simulations
        self.batch_s = batch_s
        self.dev = device
        self.latent_actions = []
        self._engine = None

    def _get_engine(self):
        n = int(self.n_simulations)  # mutated from outside by acting_ablations.py:79
        if self._engine is None or self._engine.store.n_records < n + 1:
            self._engine = BatchedMCTS(self.discount, self.root_dirichlet_alpha, n, 1, "cuda",
                                       self.root_exploration_eps, max_simulations=max(n, 64))
        e = self._engine
        e.n_simulations, e.discount = n, float(self.discount)
        e.root_dirichlet_alpha, e.root_exploration_eps = self.root_dirichlet_alpha, self.root_exploration_eps
        return e

    def run_mcts(self, state, network, temperature, deterministic):
        """:34-126 -> (action int, pi_prob np.float64[6], root_node.Q float)."""
        if not 0.0 <= temperature <= 1.0:  # :163-166 (raised before the search here, after it there)
            raise ValueError(f"Expect `temperature` to be in the range [0.0, 1.0], got {temperature}")
        eng = self._get_engine()
        obs = torch.from_numpy(np.asarray(state)).to("cuda", dtype=torch.float32).reshape(1, -1).contiguous()
        # MinMaxStats may have been replaced / edited by the caller: the host object is the truth
        eng.store.minmax.copy_(torch.tensor([[self.min_max_stats.minimum, self.min_max_stats.maximum]],
                                            dtype=torch.float64))
        use_noise = (not deterministic) and self.root_dirichlet_alpha > 0.0 and self.root_exploration_eps > 0.0
        noise = None
        if use_noise:  # :149 — same global np.random draw as the reference
            noise = np.random.dirichlet(np.ones(6, dtype=np.float32) * self.root_dirichlet_alpha)[None, :]
        u = None if deterministic else np.array([np.random.random_sample()])  # the draw of np.random.choice, :120
        action, pi, root_q, _ = eng.run_mcts(network.packed(), obs=obs, temperature=temperature,
                                             deterministic=deterministic, noise=noise, uniforms=u)
        mm = eng.store.minmax.cpu().numpy()
        self.min_max_stats.minimum, self.min_max_stats.maximum = float(mm[0, 0]), float(mm[0, 1])
        self._last_records = None
        return int(action.cpu().item()), pi[0].cpu().numpy(), float(root_q.cpu().item())

    def root_node(self):
        """Node view of the last search's root (the reference keeps `root_node` local to run_mcts)."""
        from .node import Node

        return Node.root_of(self._engine)

    def return_latent_actions(self):
        """:128-130 — actions along the path of the LAST simulation, as LongTensor[1] each."""
        eng = self._engine
        if eng is None or eng.n_simulations == 0:
            return []
        rec = eng.store.records()
        e, path = eng.n_simulations, []  # the last simulation expanded record n_simulations
        while e != 0:
            path.append(int(rec["parent_action"][0, e]))
            e = int(rec["parent"][0, e])
        self.latent_actions = [torch.tensor([a], dtype=torch.long, device=self.dev) for a in reversed(path)]
        return self.latent_actions

    def add_dirichlet_noise(self, prob, eps=0.25, alpha=0.25):
        """:132-152 (host arithmetic on 6 numbers; the batched engine mixes on device)."""
        if not isinstance(prob, np.ndarray) or prob.dtype not in (np.float32, np.float64):
            raise ValueError(f"Expect `prob` to be a numpy.array, got {prob}")
        alphas = np.ones_like(prob) * alpha
        noise = np.random.dirichlet(alphas)
        return (1 - eps) * prob + eps * noise

    def generate_play_policy(self, visits_count, temperature):
        """:154-176."""
        if not 0.0 <= temperature <= 1.0:
            raise ValueError(f"Expect `temperature` to be in the range [0.0, 1.0], got {temperature}")
        visits_count = np.asarray(visits_count, dtype=np.int64)
        if temperature > 0.0:
            exp = max(1.0, min(5.0, 1.0 / temperature))
            visits_count = np.power(visits_count, exp)
        return visits_count / np.sum(visits_count)
