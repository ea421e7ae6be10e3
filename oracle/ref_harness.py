"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference from /root/reference.

Usable only in the dev container (the GPU box has no /root/reference); it is imported by
``gen_golden.py`` and nothing else.  Nothing is copied: the reference modules are imported
from where they lie, behind stub ``matplotlib`` / ``seaborn`` modules (``utils.py:3-5`` of the
reference imports them at module top and they are not installed here).

The three sanctioned parity hooks (BASELINE.json north_star) are applied from OUTSIDE:
  1. lowest-index tie-break instead of the uniform random one at MCTS/node.py:86,
  2. the Dirichlet draw of MCTS/mcts.py:149 supplied as an input,
  3. the sampling uniform of MCTS/mcts.py:120 supplied as an input.
"""
from __future__ import annotations

import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def import_reference():
    """Returns a namespace with the reference's TowersOfHanoi, hanoi_solver, MCTS, Node,
    MinMaxStats, MuZeroNet, oneHot_encoding, compute_n_step_returns."""
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import env.hanoi as ref_env
    import env.hanoi_utils as ref_env_utils
    import MCTS.mcts as ref_mcts
    import MCTS.node as ref_node
    import MCTS.utils_mcts as ref_mm
    import networks as ref_net
    import utils as ref_utils

    assert ref_env.__file__.startswith(REFERENCE_ROOT), ref_env.__file__
    return types.SimpleNamespace(
        TowersOfHanoi=ref_env.TowersOfHanoi,
        hanoi_solver=ref_env_utils.hanoi_solver,
        MCTS=ref_mcts.MCTS,
        mcts_module=ref_mcts,
        Node=ref_node.Node,
        node_module=ref_node,
        MinMaxStats=ref_mm.MinMaxStats,
        MuZeroNet=ref_net.MuZeroNet,
        oneHot_encoding=ref_utils.oneHot_encoding,
        compute_n_step_returns=ref_utils.compute_n_step_returns,
        compute_MCreturns=ref_utils.compute_MCreturns,
        adjust_temperature=ref_utils.adjust_temperature,
    )


SELECT_LOG: list = []  # every child index chosen since the last recurrent_inference call


def _first_max_best_child(self, config, min_max_stats):
    """Hook 1: same scores as MCTS/node.py:83, first maximum instead of a random one."""
    if not self.is_expanded:
        raise ValueError("Expand leaf node first.")
    scores = self.child_Q(config, min_max_stats) + self.child_U(config)
    best = int(np.argmax(scores))
    SELECT_LOG.append(best)
    return self.children[best]


class _RandomProxy:
    def __init__(self, real, uniforms):
        self._real, self._uniforms = real, uniforms

    def choice(self, a, p=None):
        # Hook 3: legacy RandomState.choice(p=...) draws exactly one random_sample();
        # reproduce its cdf/searchsorted arithmetic with the supplied uniform.
        u = self._uniforms.pop(0)
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        return a[int(cdf.searchsorted(u, side="right"))]

    def __getattr__(self, name):
        return getattr(self._real, name)


class _NumpyProxy:
    def __init__(self, real, uniforms):
        self._real = real
        self.random = _RandomProxy(real.random, uniforms)

    def __getattr__(self, name):
        return getattr(self._real, name)


class TracingNet:
    """Delegates to a reference MuZeroNet and records every recurrent_inference output."""

    def __init__(self, net):
        self._net = net
        self.num_actions = net.num_actions
        self.calls = []

    def initial_inference(self, x):
        out = self._net.initial_inference(x)
        self.root = out
        self.calls = []
        SELECT_LOG.clear()
        return out

    def recurrent_inference(self, h, a):
        out = self._net.recurrent_inference(h, a)
        path = list(SELECT_LOG)
        SELECT_LOG.clear()
        assert path and path[-1] == int(a.argmax())
        self.calls.append((path, out))
        return out


class HookedSearch:
    """Context manager that installs the three hooks on the imported reference modules."""

    def __init__(self, ref, noises=None, uniforms=None):
        self.ref, self.noises, self.uniforms = ref, list(noises or []), list(uniforms or [])

    def __enter__(self):
        ref = self.ref
        self._saved = (ref.Node.best_child, ref.MCTS.add_dirichlet_noise, ref.mcts_module.np)
        ref.Node.best_child = _first_max_best_child
        noises = self.noises

        def add_noise(mcts_self, prob, eps=0.25, alpha=0.25):
            # Hook 2: the arithmetic of MCTS/mcts.py:150 with the draw supplied.
            if not isinstance(prob, np.ndarray) or prob.dtype not in (np.float32, np.float64):
                raise ValueError(f"Expect `prob` to be a numpy.array, got {prob}")
            return (1 - eps) * prob + eps * noises.pop(0)

        ref.MCTS.add_dirichlet_noise = add_noise
        ref.mcts_module.np = _NumpyProxy(np, self.uniforms)
        return self

    def __exit__(self, *exc):
        ref = self.ref
        ref.Node.best_child, ref.MCTS.add_dirichlet_noise, ref.mcts_module.np = self._saved
        return False


def load_reference_net(ref, n_disks, state_dict_np):
    import torch

    net = ref.MuZeroNet(rpr_input_s=3 * n_disks, action_s=6, lr=0.002, device="cpu", TD_return=True)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state_dict_np.items()})
    return net
