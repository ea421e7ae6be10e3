"""Drop-in for the reference's env/hanoi.py: TowersOfHanoi with the same constructor,
attributes, methods, return types and error behaviour (file:line cites are into the
reference), as a B=1 view over the libhmz env kernels."""
import numpy as np
import torch

from .. import _lib
from ..engine import MOVES, check_env_shape, index_state, pack_state, state_index, unpack_state

_REWARD_ILLEGAL = -100 / 1000  # env/hanoi.py:72 — python double, what callers compare against


class StateSpace:
    """Lazy stand-in for ``list(itertools.product(range(3), repeat=N))`` (env/hanoi.py:23-25):
    same indexing, ``index()``, ``len()`` and iteration without enumerating 3^N tuples."""

    def __init__(self, n_disks):
        self.n = n_disks

    def __len__(self):
        return 3 ** self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError("list index out of range")
        return index_state(i, self.n)

    def index(self, state):
        state = tuple(state)
        if len(state) != self.n or any(p not in (0, 1, 2) for p in state):
            raise ValueError(f"{state} is not in list")
        return state_index(state)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __contains__(self, state):
        try:
            self.index(state)
            return True
        except (ValueError, TypeError):
            return False


class TowersOfHanoi:
    """reference env/hanoi.py:11-151."""

    def __init__(self, N, max_steps, init_state_idx=0, goal_peg=2):
        check_env_shape(N, max_steps)
        self.discs = N
        self.n_pegs = 3
        self.states = StateSpace(N)
        self.oneH_s_size = N * 3
        self.goal = tuple([goal_peg] * N)
        self.init_state_idx = init_state_idx
        self.moves = list(MOVES)
        self.max_steps = max_steps
        self.reset_check = False
        self.step_counter = 0
        self._goal_peg = goal_peg
        self._lib = None

    # -- device plumbing (lazy so that constructing the object needs no GPU) -------------
    def _dev(self):
        if self._lib is None:
            _lib.require_cuda()
            self._lib = _lib.load()
            self._word = torch.zeros(4, dtype=torch.int32, device="cuda")  # [word, obs_word, -, -]
            self._act = torch.zeros(4, dtype=torch.uint8, device="cuda")
            self._rwd = torch.zeros(4, dtype=torch.float32, device="cuda")
            self._flag = torch.zeros(4, dtype=torch.uint8, device="cuda")
            self._obs = torch.zeros(3 * self.discs, dtype=torch.float32, device="cuda")
        return self._lib

    def _word_now(self):
        return pack_state(self.c_state) | (int(self.step_counter) << (2 * self.discs))

    def _onehot(self, state):
        lib = self._dev()
        self._word[1] = pack_state(state)
        _lib.check(lib.hmz_env_onehot(_lib.ptr(self._word[1:]), _lib.ptr(self._obs), 1, self.discs,
                                      _lib.current_stream()))
        return self._obs.cpu().numpy().astype(np.float64)

    # -- reference surface ------------------------------------------------------------------
    def step(self, action):
        assert self.reset_check, "Need to reset env before taking a step"  # env/hanoi.py:49
        move = self.moves[action]  # IndexError for actions outside 0..5, as in the reference
        del move
        lib = self._dev()
        st = _lib.current_stream()
        self._word[0] = self._word_now()
        self._act[0] = int(action)
        _lib.check(lib.hmz_env_step(_lib.ptr(self._word), _lib.ptr(self._act), _lib.ptr(self._rwd),
                                    _lib.ptr(self._flag), _lib.ptr(self._word[1:]), 1, self.discs, self.max_steps,
                                    self._goal_peg, 0, 0, st))
        _lib.check(lib.hmz_env_onehot(_lib.ptr(self._word[1:]), _lib.ptr(self._obs), 1, self.discs, st))
        word = int(self._word[0].item()) & 0xFFFFFFFF
        flags = int(self._flag[0].item())
        shift = 2 * self.discs
        self.c_state = unpack_state(word & ((1 << shift) - 1), self.discs)
        self.step_counter = word >> shift
        done = bool(flags & _lib.FLAG_DONE)
        illegal_move = bool(flags & _lib.FLAG_ILLEGAL)
        if done:
            self.reset_check = False
        if flags & _lib.FLAG_GOAL:
            rwd = 100
        elif illegal_move:
            rwd = _REWARD_ILLEGAL
        else:
            rwd = 0
        return self._obs.cpu().numpy().astype(np.float64), rwd, done, illegal_move

    def reset(self):
        self.reset_check = True
        self.c_state = self.states[self.init_state_idx]
        self.oneH_c_state = self._onehot(self.c_state)
        return self.oneH_c_state

    def random_reset(self):
        self.reset_check = True
        while True:  # env/hanoi.py:105-109: same np.random stream consumption as the reference
            random_indx = np.random.randint(len(self.states))
            self.c_state = self.states[random_indx]
            if self.c_state != self.goal:
                break
        self.oneH_c_state = self._onehot(self.c_state)
        return self.oneH_c_state

    def current_state(self):
        return list(self.c_state)

    def _discs_on_peg(self, peg):
        return [disc for disc in range(self.discs) if self.c_state[disc] == peg]

    def _legal_bits(self):
        lib = self._dev()
        self._word[2] = pack_state(self.c_state)
        _lib.check(lib.hmz_env_legal_mask(_lib.ptr(self._word[2:]), _lib.ptr(self._flag[1:]), 1, self.discs,
                                          _lib.current_stream()))
        return int(self._flag[1].item())

    def _move_allowed(self, move):
        return bool((self._legal_bits() >> self.moves.index(tuple(move))) & 1)

    def _get_moved_state(self, move):
        a = self.moves.index(tuple(move))
        if not (self._legal_bits() >> a) & 1:
            # the reference leaves `disc_to_move` unbound for a disallowed move (env/hanoi.py:142-148)
            raise UnboundLocalError("cannot access local variable 'disc_to_move' where it is not associated with a value")
        lib = self._dev()
        self._word[2] = pack_state(self.c_state)  # counter 0: never truncates for max_steps >= 2
        self._act[1] = a
        _lib.check(lib.hmz_env_step(_lib.ptr(self._word[2:]), _lib.ptr(self._act[1:]), _lib.ptr(self._rwd[1:]),
                                    _lib.ptr(self._flag[1:]), _lib.ptr(self._word[3:]), 1, self.discs,
                                    max(self.max_steps, 2), self._goal_peg, 0, 0, _lib.current_stream()))
        return unpack_state(int(self._word[3].item()), self.discs)
