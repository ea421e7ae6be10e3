"""Drop-in for the hot-path half of the reference's utils.py (oneHot_encoding :9-25,
adjust_temperature :89-96).  Plotting / logging helpers of the reference are out of scope."""
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
