"""Drop-in for the reference's MCTS/node.py.  In this engine the tree lives in the device tree
store (include/hmz.h: one 128-byte record per expanded node), so a ``Node`` is a VIEW
``(search, record | parent slot)`` onto that store with the reference's attribute and method
names (cites are reference MCTS/node.py:line).  Expansion and backup are performed by the search
kernels (hmz_search_run); the view's ``expand`` / ``backup`` therefore only reproduce the reference's
error behaviour."""
import ctypes as C

import numpy as np
import torch

from .. import _lib

NO_CHILD = _lib.NO_CHILD


class Node:
    def __init__(self, prior=None, move=None, parent=None, _engine=None, _record=None):
        """Reference signature (:9-28).  Views are created by ``Node.root_of(engine)``; a Node built
        with the reference signature alone is a detached, unexpanded node."""
        self.prior, self.move, self.parent = prior, move, parent
        self._engine, self._record = _engine, _record  # record index when expanded, else None
        self._snap = None

    # ---- construction of views ---------------------------------------------------------------
    @classmethod
    def root_of(cls, engine, search=0):
        """View of the root of `engine`'s (BatchedMCTS) search number `search` after a run."""
        node = cls(prior=0.0, move=None, parent=None, _engine=engine, _record=0)  # root prior = 0.0 (mcts.py:52-54)
        node._search = search
        return node

    def _rec(self):
        if self._snap is None:
            self._snap = self._engine.store.records()
        return self._snap

    # ---- reference fields ----------------------------------------------------------------------
    @property
    def is_expanded(self):
        return self._engine is not None and self._record is not None

    def _slot(self, name):
        p = self.parent
        return p._rec()[name][p._search, p._record, self.move]

    @property
    def N(self):
        if self._engine is None:
            return 0
        if self.parent is None:
            return int(self._rec()["N"][self._search, 0].sum())
        return int(self._slot("N"))

    @property
    def W(self):
        if self._engine is None:
            return 0.0
        if self.parent is None:
            return float(self._engine.store.root_W[self._search].item())
        return float(self._slot("W"))

    @property
    def rwd(self):
        if self._engine is None or self.parent is None:
            return 0.0
        return float(self._slot("rwd"))

    @property
    def h_state(self):
        if not self.is_expanded:
            return None
        return self._engine.store.latents[self._search, self._record].float().cpu().numpy()

    @property
    def children(self):
        if not self.is_expanded:
            return []
        rec = self._rec()
        out = []
        for a in range(6):
            child_rec = int(rec["child"][self._search, self._record, a])
            prior = rec["prior"][self._search, self._record, a]
            if self._record == 0 and self._engine.store.desc.root_prior_is_f64:
                prior = np.float64(self._engine.store.root_prior[self._search, a].item())
            c = Node(prior=prior, move=a, parent=self, _engine=self._engine,
                     _record=None if child_rec == NO_CHILD else child_rec)
            c._search, c._snap = self._search, rec
            out.append(c)
        return out

    # ---- reference methods ---------------------------------------------------------------------
    def expand(self, prior, h_state, reward):
        if self.is_expanded:
            raise RuntimeError("Node has already been expanded")  # :40-41
        raise NotImplementedError("expansion is performed on the device by the search kernels (hmz_search_run); "
                                  "Node is a view onto the tree store")

    def backup(self, value, config, min_max_stats):
        raise NotImplementedError("backup is performed on the device by the search kernels (hmz_search_run); "
                                  "Node is a view onto the tree store")

    def _scores(self, config):
        if not self.is_expanded:
            raise ValueError("Expand leaf node first.")  # :80-81
        eng, lib = self._engine, self._engine.lib
        B = eng.B
        rec = torch.zeros(B, dtype=torch.int16, device=eng.device)
        rec[self._search] = self._record
        nn = torch.zeros(B, dtype=torch.int32, device=eng.device)
        nn[self._search] = self.N
        q = torch.empty(B, 6, dtype=torch.float32, device=eng.device)
        u = torch.empty(B, 6, dtype=torch.float32, device=eng.device)
        best = torch.empty(B, dtype=torch.int32, device=eng.device)
        _lib.check(lib.hmz_search_child_scores(C.byref(eng.store.desc), _lib.ptr(rec), _lib.ptr(nn), _lib.ptr(eng._table),
                                               float(config.discount), _lib.ptr(q), _lib.ptr(u), _lib.ptr(best),
                                               _lib.current_stream()))
        return q[self._search].cpu().numpy(), u[self._search].cpu().numpy(), int(best[self._search].item())

    def best_child(self, config, min_max_stats):
        """:72-88 with the lowest-index tie-break (the documented replacement of the random one)."""
        _, _, best = self._scores(config)
        return self.children[best]

    def child_Q(self, config, min_max_stats):
        return self._scores(config)[0]

    def child_U(self, config):
        return self._scores(config)[1]

    @property
    def Q(self):
        n = self.N
        return 0.0 if n == 0 else self.W / n  # :125-131

    @property
    def child_N(self):
        return np.array([child.N for child in self.children], dtype=np.int32)  # :133-136

    @property
    def has_parent(self):
        return isinstance(self.parent, Node)
