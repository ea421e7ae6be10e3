"""GPU, system level: the whole loop of the reference's training_main.py — batched self-play, episode
post-processing, replay ring, learner step — runs on the device and LEARNS 3-disk Tower of Hanoi: from a torch
default-initialised network the mean episode length falls from ~max_steps to near the optimal 7 moves."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("acting_mode", [0, 2])  # acting on the FFMA kernel / on the tensor cores (HMZ_MODE_FP32X3)
def test_batched_training_loop_learns_three_disk_hanoi(acting_mode):
    from muzero_hanoi_b200.networks import MuZeroNet
    from muzero_hanoi_b200.trainer import BatchedMuzero

    torch.manual_seed(1)
    np.random.seed(1)
    net = MuZeroNet(9, 6, 0.002, "cpu", TD_return=True)
    mz = BatchedMuzero(net.state_dict(), 3, 200, 512, n_mcts_simulations=25, n_update_x_loop=8, acting_mode=acting_mode, seed=1)
    hist = mz.training_loop(160, min_replay_size=5000)
    lens = [h[1] for h in hist if h[1] == h[1]]
    early, late = float(np.mean(lens[:15])), float(np.mean(lens[-15:]))
    assert mz.learner.step_index > 500 and len(mz.buffer) > 5000
    assert early > 60 and late < 20, (early, late)  # optimal 7; sampled actions (T = 1) and Dirichlet noise keep it above
    # the trained weights go back into the reference-shaped module unchanged in form
    net.load_state_dict({k: v.cpu() for k, v in mz.learner.state_dict().items()})
    assert all(torch.isfinite(p).all() for p in net.parameters())
