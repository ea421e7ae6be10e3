"""TEST INFRASTRUCTURE ONLY — ctypes loader of oracle/c/libhmz_oracle.so (the C restatement of
the env step and tree arithmetic).  Imported by tests/, smoke() and bench.py's CPU baseline only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "libhmz_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", os.path.join(HERE, "c")], check=True)
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def env_step(words, actions, n_disks, max_steps, goal_peg=2, auto_reset=False, reset_word=0):
    """words uint32 [B] (modified in place), actions uint8 [B] -> (rewards f32, flags u8, obs_words u32)."""
    B = len(words)
    rewards, flags, obs = np.zeros(B, np.float32), np.zeros(B, np.uint8), np.zeros(B, np.uint32)
    load().oracle_env_step(_p(words), _p(actions), _p(rewards), _p(flags), _p(obs), C.c_int64(B), n_disks, max_steps,
                           goal_peg, int(auto_reset), C.c_uint32(reset_word))
    return rewards, flags, obs


def legal_mask(words, n_disks):
    out = np.zeros(len(words), np.uint8)
    load().oracle_legal_mask(_p(words), _p(out), C.c_int64(len(words)), n_disks)
    return out


def solver(words, n_disks, goal_peg=2):
    out = np.zeros(len(words), np.uint32)
    load().oracle_solver(_p(words), _p(out), C.c_int64(len(words)), n_disks, goal_peg)
    return out


def search_injected(prior, prior_is_f64, minmax, r, p, v, discount, table, want_depth=False):
    """prior f64 [B,6]; minmax f64 [B,2] (updated in place); r, v f32 [S,B]; p f32 [S,B,6].
    Returns (visits int32 [B,6], root_q f64 [B], leaf_depth uint16 [S,B] | None)."""
    S, B = r.shape
    prior = np.ascontiguousarray(prior, np.float64)
    r, v, p = (np.ascontiguousarray(x, np.float32) for x in (r, v, p))
    assert minmax.dtype == np.float64 and minmax.flags.c_contiguous and table.dtype == np.float64 and len(table) > S
    visits, root_q = np.zeros((B, 6), np.int32), np.zeros(B, np.float64)
    depth = np.zeros((S, B), np.uint16) if want_depth else None
    load().oracle_search_injected(C.c_int64(B), S, C.c_double(discount), _p(prior), int(bool(prior_is_f64)), _p(minmax),
                                  _p(r), _p(p), _p(v), _p(table), _p(visits), _p(root_q), _p(depth))
    return visits, root_q, depth
