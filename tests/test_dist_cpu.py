"""CPU (gloo, world_size 2): the multi-GPU plumbing — game sharding and the all-gather of move
records (the only collective on the path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_slot(rank, B):
    g = torch.Generator().manual_seed(100 + rank)
    return dict(state=torch.randint(0, 1 << 20, (B,), dtype=torch.int32, generator=g),
                action=torch.randint(0, 6, (B,), dtype=torch.uint8, generator=g),
                reward=torch.rand(B, generator=g), flags=torch.randint(0, 16, (B,), dtype=torch.uint8, generator=g),
                visits=torch.randint(0, 101, (B, 6), dtype=torch.int16, generator=g),
                root_q=torch.randn(B, generator=g, dtype=torch.float64),
                game_lo=(torch.arange(B) + rank * B).to(torch.int16))


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from muzero_hanoi_b200 import dist as hdist

    r, w, _ = hdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    gathered = hdist.all_gather_records(hdist.pack_records(_fake_slot(rank, B)))
    assert gathered.shape == (world * B, 32)
    for src in range(world):
        want = _fake_slot(src, B)
        got = hdist.unpack_records(gathered[src * B:(src + 1) * B])
        for k in hdist.RECORD_FIELDS:
            assert torch.equal(got[k], want[k]), (src, k)
    # the pipelined form the bench uses: three moves through a two-deep RecordGather
    pipe = hdist.RecordGather(B, world, "cpu", depth=2)
    for move in range(3):
        i = pipe.submit(hdist.pack_records(_fake_slot(10 * move + rank, B)))
        got = hdist.unpack_records(pipe.result(i))
        for src in range(world):
            want = _fake_slot(10 * move + src, B)
            for k in hdist.RECORD_FIELDS:
                assert torch.equal(got[k][src * B:(src + 1) * B], want[k]), (move, src, k)
    lo, hi = hdist.shard_range(10, rank, world)
    out[rank] = (lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_records_world2():
    world, B = 2, 37
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, out), nprocs=world, join=True)
    assert dict(out) == {0: (0, 5), 1: (5, 10)}


def test_shard_range_covers_everything():
    from muzero_hanoi_b200 import dist as hdist

    for n in (0, 1, 7, 65536):
        for world in (1, 2, 3, 8):
            spans = [hdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_record_pack_roundtrip_single_process():
    from muzero_hanoi_b200 import dist as hdist

    slot = _fake_slot(0, 129)
    buf = hdist.all_gather_records(hdist.pack_records(slot))  # world 1: the wire records themselves
    assert buf.dtype == torch.uint8 and buf.shape == (129, 32)
    back = hdist.unpack_records(buf)
    for k in hdist.RECORD_FIELDS:
        assert torch.equal(back[k], slot[k])


def test_wire_record_layout_matches_the_c_struct():
    """dist's byte offsets against the ctypes mirror of hmz_move_record_t (itself checked against the header by gcc)."""
    import numpy as np

    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200 import dist as hdist

    for name, (lo, hi) in hdist._OFF.items():
        f = getattr(_lib.MoveRecord, name)
        assert (f.offset, f.offset + f.size) == (lo, hi), name
    assert np.dtype(_lib.RECORD_DTYPE).itemsize == hdist.RECORD_BYTES == _lib.RECORD_BYTES
