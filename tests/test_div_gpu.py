"""GPU: the search kernels' exact-division shortcut (reciprocal table / __drcp_rn + two FMA corrections,
hmz_tree.cuh) must equal IEEE float64 division bit for bit — it sits on the bit-exact visit-count path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_exact_division_shortcut_matches_ieee_division(seed):
    from muzero_hanoi_b200 import _lib

    lib = _lib.load()
    counters = torch.zeros(2, dtype=torch.int64, device="cuda")
    n = 1 << 28  # 2.7e8 operand pairs per seed and per form
    _lib.check(lib.hmz_debug_div_check(n, seed, _lib.ptr(counters), _lib.current_stream()))
    torch.cuda.synchronize()
    assert counters.tolist() == [0, 0], f"mismatches (by count, by range) = {counters.tolist()} of {n}"
