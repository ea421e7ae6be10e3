"""Drop-in for the acting-path half of the reference's utils.py (oneHot_encoding :9-25,
compute_n_step_returns :28-72, compute_MCreturns :75-86, adjust_temperature :89-96).  Plotting / logging helpers of the reference
are out of scope."""
import numpy as np


def oneHot_encoding(x, n_integers):
    """float64 vector of length len(x)*n_integers, disk-major (reference utils.py:9-25).
    Host-side formatting of a handful of integers; the batched device form is
    ``hmz_env_onehot`` (engine.VecHanoi.onehot)."""
    x = np.asarray(x, dtype=np.int64)
    out = np.zeros((x.shape[0], n_integers))
    out[np.arange(x.shape[0]), x] = 1
    return out.reshape(-1)


def adjust_temperature(episode):
    """Temperature schedule of self-play (reference utils.py:89-96)."""
    if episode < 500:
        return 1.0
    if episode < 750:
        return 0.5
    return 0.1


def compute_n_step_returns(rwds, root_values, n_step, discount):
    """n-step TD targets of one episode (reference utils.py:28-72), computed by the device kernel of
    replay.EpisodeStore (the batched form handles every finished game of a self-play batch at once).
    Rewards must be the env's own values 0 / 100 / -100/1000 (env/hanoi.py:62,66,72)."""
    import torch

    from .replay import EpisodeStore

    assert n_step > 0, "the n_step return must be greater than zero"
    assert len(rwds) == len(root_values), "`rewards` and `root_values` don have the same length."
    T = len(rwds)
    if T == 0:
        return []
    codes = {0: 0, 100: 4 | 1, -100 / 1000: 2}
    try:
        flags = [codes[r] for r in rwds]
    except KeyError as e:
        raise ValueError(f"reward {e.args[0]!r} is not one the Hanoi env produces") from None
    st = EpisodeStore(1, T, 1)
    st.flags[:, 0] = torch.tensor(flags, dtype=torch.uint8)
    st.root_q[:, 0] = torch.tensor([float(q) for q in root_values], dtype=torch.float64)
    st.ep_len.fill_(T)
    ret, _ = st.post_process(n_step, discount)
    return ret[:, 0].cpu().tolist()


def compute_MCreturns(rwds, discount):
    """Discounted reward-to-go of one episode (reference utils.py:75-86), on the device kernel."""
    import torch

    from .replay import EpisodeStore

    T = len(rwds)
    if T == 0:
        return []
    codes = {0: 0, 100: 4 | 1, -100 / 1000: 2}
    try:
        flags = [codes[r] for r in rwds]
    except KeyError as e:
        raise ValueError(f"reward {e.args[0]!r} is not one the Hanoi env produces") from None
    st = EpisodeStore(1, T, 1)
    st.flags[:, 0] = torch.tensor(flags, dtype=torch.uint8)
    st.ep_len.fill_(T)
    ret, _ = st.post_process_mc(discount)
    return ret[:, 0].cpu().tolist()
