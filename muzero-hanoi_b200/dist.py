"""Multi-GPU plumbing: one process per GPU (torchrun), games sharded by rank with no data-path
collective; the ONLY exchange is the all-gather of finished-move trajectory records towards the
replay buffer.  Works over NCCL on GPUs and over gloo on CPU (tests).

Wire format = the trajectory ring's own element, ``hmz_move_record_t`` (include/hmz.h), 32 bytes per game-move, written by
the fused end-of-move kernel: nothing is repacked between the kernel and the collective."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

RECORD_BYTES = 32
RECORD_FIELDS = ("root_q", "state", "reward", "visits", "action", "flags", "game_lo")
# byte offsets inside a record: root_q f64 | state u32 | reward f32 | visits 6 x u16 | action u8 | flags u8 | game_lo u16
_OFF = dict(root_q=(0, 8), state=(8, 12), reward=(12, 16), visits=(16, 28), action=(28, 29), flags=(29, 30), game_lo=(30, 32))
_DT = dict(root_q=torch.float64, state=torch.int32, reward=torch.float32, visits=torch.int16, action=torch.uint8,
           flags=torch.uint8, game_lo=torch.int16)


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of games owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(fields: dict) -> torch.Tensor:
    """Host-side construction of wire records from per-field tensors (tests, tooling; the engine's kernel writes the
    records itself): uint8 [B, 32]."""
    B = fields["state"].shape[0]
    out = torch.zeros(B, RECORD_BYTES, dtype=torch.uint8, device=fields["state"].device)
    for k, (lo, hi) in _OFF.items():
        if k in fields:
            out[:, lo:hi] = fields[k].to(_DT[k]).contiguous().view(torch.uint8).reshape(B, hi - lo)
    return out


def unpack_records(buf: torch.Tensor) -> dict:
    """uint8 [n, 32] wire records -> dict of per-field tensors (visits int16 [n, 6])."""
    n = buf.shape[0]
    out = {}
    for k, (lo, hi) in _OFF.items():
        v = buf[:, lo:hi].contiguous().view(_DT[k])
        out[k] = v.reshape(n, 6) if k == "visits" else v.reshape(n)
    return out


def all_gather_records(records: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Blocking all-gather of one move's records (uint8 [B, 32]) from every rank -> uint8 [world * B, 32], rank-major.
    Every rank must contribute the same B."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return records
    world = dist.get_world_size()
    if out is None:
        out = torch.empty(world * records.shape[0], RECORD_BYTES, dtype=torch.uint8, device=records.device)
    dist.all_gather_into_tensor(out.view(-1), records.contiguous().view(-1))
    return out


class RecordGather:
    """The per-move all-gather, off the compute stream: every ``submit`` enqueues ``all_gather_into_tensor`` of the move's
    records on a SIDE stream into one of ``depth`` receive buffers, so the collective overlaps the next move's search
    ("plain NCCL on a side stream").  Ordering: the side stream waits for the event recorded after the
    move on the compute stream; the compute stream waits, ``depth`` submits later, for the gather that read the ring
    slot / receive buffer about to be reused — so the trajectory ring needs more than ``depth`` slots."""

    def __init__(self, B, world, device, depth=2):
        self.B, self.world, self.depth = int(B), int(world), int(depth)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.out = torch.empty(self.depth, self.world * self.B, RECORD_BYTES, dtype=torch.uint8, device=self.device)
        self.count = 0
        if self.cuda:
            self.stream = torch.cuda.Stream(self.device)
            self.done = [torch.cuda.Event() for _ in range(self.depth)]
            self.moved = [torch.cuda.Event() for _ in range(self.depth)]

    def submit(self, records: torch.Tensor) -> int:
        """records: uint8 [B, 32], produced on the current stream.  Returns the receive-buffer index."""
        i = self.count % self.depth
        active = dist.is_initialized() and dist.get_world_size() > 1
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            if self.count >= self.depth:
                cur.wait_event(self.done[i])  # the gather that used this buffer (and its ring slot) has completed
            self.moved[i].record(cur)
            self.stream.wait_event(self.moved[i])
            with torch.cuda.stream(self.stream):
                if active:
                    dist.all_gather_into_tensor(self.out[i].view(-1), records.view(-1))
                else:
                    self.out[i, : self.B].copy_(records, non_blocking=True)
                self.done[i].record(self.stream)
        elif active:
            dist.all_gather_into_tensor(self.out[i].view(-1), records.contiguous().view(-1))
        else:
            self.out[i, : self.B].copy_(records)
        self.count += 1
        return i

    def result(self, i) -> torch.Tensor:
        """Receive buffer ``i`` (uint8 [world * B, 32]) once its gather is done, ordered into the current stream."""
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_event(self.done[i])
        return self.out[i]

    def drain(self):
        if self.cuda:
            self.stream.synchronize()
