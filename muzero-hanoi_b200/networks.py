"""Drop-in for the reference's networks.py: MuZeroNet with the same constructor, sub-module
names, state_dict keys and method signatures (cites are reference networks.py:line).

The ACTING methods — initial_inference and recurrent_inference, the only ones on the self-play
hot path (MCTS/mcts.py:50,102 of the reference) — run on the libhmz kernels.  represent /
dynamics / prediction / update keep their torch-module form for the reference's gradient analyses; the
training step itself also exists on the device (learner.Learner)."""
import numpy as np
import torch
import torch.nn as nn
import torch.optim as opt

from . import _lib
from .engine import PackedWeights


class MuZeroNet(nn.Module):
    def __init__(self, rpr_input_s, action_s, lr, device, reward_s=1, h1_s=256, reprs_output_size=64,
                 weight_decay=1e-4, TD_return=False):
        super().__init__()
        self.dev = device
        self.num_actions = action_s
        self.TD_return = TD_return
        self.reprs_output_size = reprs_output_size
        self.support_size = 33 if TD_return else 1  # networks.py:33-37

        def mlp(i, o):
            return nn.Sequential(nn.Linear(i, h1_s), nn.ReLU(), nn.Linear(h1_s, o))

        self.representation_net = mlp(rpr_input_s, reprs_output_size)
        self.dynamic_net = mlp(reprs_output_size + action_s, reprs_output_size)
        self.rwd_net = mlp(reprs_output_size, self.support_size)
        self.policy_net = mlp(reprs_output_size, action_s)
        self.value_net = mlp(reprs_output_size, self.support_size)
        self.optimiser = opt.Adam(self.parameters(), lr)
        self._rpr_input_s = rpr_input_s
        self._packed = {}  # mode -> (parameter version stamp, PackedWeights)

    # ------------------------------------------------------------------ kernel plumbing
    def _stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def packed(self, mode=_lib.MODE_FP32):
        """PackedWeights of the current parameters; repacked whenever a parameter changed
        (optimiser step, load_state_dict, reset_param lesion ...)."""
        if not self.TD_return or self.reprs_output_size != 64 or self._rpr_input_s % 3 != 0:
            raise NotImplementedError("libhmz kernels cover the reference configuration: TD_return=True (support 33), "
                                      "h1_s=256, reprs_output_size=64, Hanoi one-hot input of width 3N")
        stamp = self._stamp()
        hit = self._packed.get(mode)
        if hit is None or hit[0] != stamp:
            hit = (stamp, PackedWeights(self.state_dict(), self._rpr_input_s // 3, mode))
            self._packed[mode] = hit
        return hit[1]

    @torch.no_grad()
    def initial_inference(self, x):
        """networks.py:71-94 -> (h np.float32[64], rwd 0.0, pi_probs np.float32[6], value float)."""
        w = self.packed()
        obs = torch.as_tensor(x, dtype=torch.float32).reshape(1, -1).to("cuda").contiguous()
        h = torch.empty(1, 64, dtype=torch.float32, device="cuda")
        p0 = torch.empty(1, 6, dtype=torch.float32, device="cuda")
        v0 = torch.empty(1, dtype=torch.float32, device="cuda")
        w.initial(1, obs=obs, latents_out=h, out_rows_per_item=1, latent_dtype=_lib.LATENT_F32, p0=p0, v0=v0)
        return h[0].cpu().numpy(), 0.0, p0[0].cpu().numpy(), v0.cpu().item()

    @torch.no_grad()
    def recurrent_inference(self, h_state, action):
        """networks.py:96-116; ``action`` is the one-hot float vector the reference passes."""
        w = self.packed()
        h_in = torch.as_tensor(h_state, dtype=torch.float32).reshape(1, 64).to("cuda").contiguous()
        a = torch.as_tensor(action).reshape(-1).argmax().to(torch.uint8).reshape(1).to("cuda")
        h = torch.empty(1, 64, dtype=torch.float32, device="cuda")
        r = torch.empty(1, dtype=torch.float32, device="cuda")
        p = torch.empty(1, 6, dtype=torch.float32, device="cuda")
        v = torch.empty(1, dtype=torch.float32, device="cuda")
        w.recurrent(1, latents_in=h_in, in_rows_per_item=1, in_row=None, actions=a, latents_out=h,
                    out_rows_per_item=1, out_row=0, latent_dtype=_lib.LATENT_F32, r=r, p=p, v=v)
        return h[0].cpu().numpy(), r.cpu().item(), p[0].cpu().numpy(), v.cpu().item()

    # ---------------------------------------------- differentiable forms (learner / analyses)
    def update(self, loss):
        self.optimiser.zero_grad()
        loss.backward()
        self.optimiser.step()

    def represent(self, x):
        return self.normalize_h_state(self.representation_net(x))

    def dynamics(self, h_state, action):
        new_h_state = self.dynamic_net(torch.cat([h_state, action], dim=-1))
        rwd_prediction = self.rwd_net(new_h_state)
        if self.TD_return:
            rwd_prediction = self.logits_to_transformed_expected_value(rwd_prediction)
        return self.normalize_h_state(new_h_state), rwd_prediction

    def prediction(self, h):
        pi_logits = self.policy_net(h)
        value_logits = self.value_net(h)
        if self.TD_return:
            value_logits = self.logits_to_transformed_expected_value(value_logits)
        return pi_logits, value_logits

    def logits_to_transformed_expected_value(self, logits):
        max_value = (self.support_size - 1) // 2
        probs = torch.softmax(logits, dim=-1)
        return self._signed_parabolic(self._transform_from_2hot(probs, -max_value, max_value))

    def _transform_from_2hot(self, probs, min_value, max_value):
        support_space = torch.linspace(min_value, max_value, self.support_size, device=probs.device)
        return torch.sum(probs * support_space.expand_as(probs), dim=-1, keepdim=True)

    def _signed_parabolic(self, x, eps=1e-3):
        z = torch.sqrt(1 + 4 * eps * (eps + 1 + torch.abs(x))) / 2 / eps - 1 / 2 / eps
        return torch.sign(x) * (torch.square(z) - 1)

    def normalize_h_state(self, h_state):
        _min = h_state.min(dim=-1, keepdim=True)[0]
        _max = h_state.max(dim=-1, keepdim=True)[0]
        return (h_state - _min) / (_max - _min + 1e-8)

    def set_pol_pertubation(self, pertub_magnitude):
        self.perturb_p_magnitude = pertub_magnitude

    def reset_param(self, l):
        """Lesion re-initialisation (networks.py:201-205): U(-k, k), k = sqrt(1/latent), weight and bias."""
        k = np.sqrt(1 / self.reprs_output_size)
        if isinstance(l, nn.Linear):
            nn.init.uniform_(l.weight, a=-k, b=k)
            nn.init.uniform_(l.bias, a=-k, b=k)
