"""TEST INFRASTRUCTURE ONLY — CPU oracle for the MuZero-Hanoi acting hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the path that is shipped or measured as ours.

Contents
--------
``port.py``        numpy/torch restatement of the reference's env, solver,
                   search and network inference (cites reference file:line).
``c/hmz_oracle.c`` plain-C restatement of the env + search (fast checker for
                   the 4,096-search parity cases), loaded through ``cport.py``.
``ref_harness.py`` imports the UNMODIFIED reference from ``/root/reference``
                   (dev container only) behind stub plotting modules.
``gen_golden.py``  runs the reference through ``ref_harness`` and writes the
                   fixtures under ``tests/golden/`` (and checks ``port.py``
                   against the reference bit-for-bit while doing so).

Pinning status: the reference ships no tests / golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against *outputs of the reference itself
executed in this container* (torch 2.11.0, numpy 2.3.5) — see
``tests/golden/MANIFEST.json`` for the versions each fixture was made with.
"""
