"""CPU: libhmz.so loads, exports every symbol include/hmz.h declares, and the ctypes mirror
matches the header (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "hmz.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmz_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from muzero_hanoi_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in hmz.h but not exported by libhmz.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_struct_layout():
    from muzero_hanoi_b200 import _lib

    assert ctypes.sizeof(_lib.NodeRecord) == 128 and ctypes.sizeof(_lib.NodeHalf) == 64 and ctypes.sizeof(_lib.ChildSlot) == 16
    assert _lib.ChildSlot.rwd.offset == 8 and _lib.ChildSlot.N.offset == 12 and _lib.ChildSlot.child.offset == 14
    assert _lib.NodeHalf.prior.offset == 48 and _lib.NodeHalf.parent.offset == 60 and _lib.NodeHalf.parent_action.offset == 62
    assert ctypes.sizeof(_lib.SearchDesc) == 80 and _lib.SearchDesc.capture.offset == 48 and _lib.SearchDesc.schedule.offset == 76


def test_ctypes_structs_match_the_header_as_compiled_by_gcc(tmp_path):
    """sizeof / offsetof of every struct in include/hmz.h, printed by a C program, against the ctypes mirror."""
    import subprocess

    from muzero_hanoi_b200 import _lib

    structs = {"hmz_move_record_t": _lib.MoveRecord, "hmz_child_t": _lib.ChildSlot, "hmz_half_t": _lib.NodeHalf, "hmz_node_t": _lib.NodeRecord,
               "hmz_search_t": _lib.SearchDesc, "hmz_selfplay_t": _lib.SelfPlayDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hmz.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} . %zu\\n", sizeof({cname}));')
        for fname, *_ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("  return 0;\n}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for row in filter(None, out):
        cname, fname, val = row.split()
        ct = structs[cname]
        want = ctypes.sizeof(ct) if fname == "." else getattr(ct, fname).offset
        assert int(val) == want, f"{cname}.{fname}: header {val}, ctypes {want}"
        seen += 1
    assert seen > 60


def test_version_and_error_string(lib):
    assert lib.hmz_version() >= 200
    assert lib.hmz_build_flags() == b"", "libhmz.so was built with non-default tuning switches"
    assert isinstance(lib.hmz_last_error(), bytes)
    assert lib.hmz_launch_count() >= 0


def test_argument_validation_without_gpu(lib):
    """Argument errors are reported before any CUDA call, so they are checkable on CPU."""
    from muzero_hanoi_b200 import _lib

    rc = lib.hmz_env_step(None, None, None, None, None, 4, 3, 200, 2, 0, 0, None)
    assert rc == _lib.ERR_INVALID
    buf = (ctypes.c_uint32 * 4)()
    rc = lib.hmz_env_step(buf, buf, buf, buf, None, 4, 13, 200, 2, 0, 0, None)
    assert rc == _lib.ERR_UNSUPPORTED and b"n_disks" in lib.hmz_last_error()
    rc = lib.hmz_env_step(buf, buf, buf, buf, None, 4, 12, 256, 2, 0, 0, None)
    assert rc == _lib.ERR_UNSUPPORTED and b"max_steps" in lib.hmz_last_error()


def test_no_cpu_fallback():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from muzero_hanoi_b200.engine import VecHanoi

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VecHanoi(3, 200, 16)
