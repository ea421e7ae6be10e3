"""muzero-hanoi_b200 — B200-native batched MuZero acting engine for Tower of Hanoi.

Drop-in for ONE hot path of A-Andrews/Muzero-Hanoi (env step -> MCTS -> g+f MLP -> root
policy -> action).  The same-named modules of the reference live here:

    muzero_hanoi_b200.env.hanoi.TowersOfHanoi     (reference env/hanoi.py)
    muzero_hanoi_b200.env.hanoi_utils.hanoi_solver (reference env/hanoi_utils.py)
    muzero_hanoi_b200.MCTS.mcts.MCTS / MCTS.node.Node / MCTS.utils_mcts.MinMaxStats
    muzero_hanoi_b200.networks.MuZeroNet
    muzero_hanoi_b200.utils.oneHot_encoding

plus the batched engine (``engine.VecHanoi``, ``engine.BatchedMCTS``, ``engine.SelfPlay``).
All compute goes through the C ABI of ``libhmz.so`` (include/hmz.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
