"""Multi-GPU plumbing: one process per GPU (torchrun), games sharded by rank with no data-path
collective; the ONLY exchange is the all-gather of finished-move trajectory records towards the
replay buffer.  Works over NCCL on GPUs and over gloo on CPU (tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

RECORD_FIELDS = ("state", "action", "reward", "flags", "visits", "root_q")


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


def pack_records(slot: dict) -> torch.Tensor:
    """One move's records of this rank as a single uint8 [B, 26] tensor (wire format:
    state u32 | root_q f32 | reward f32 | visits 6 x u16 | action u8 | flags u8)."""
    B = slot["state"].shape[0]
    out = torch.empty(B, 26, dtype=torch.uint8, device=slot["state"].device)
    out[:, 0:4] = slot["state"].contiguous().view(torch.uint8).reshape(B, 4)
    out[:, 4:8] = slot["root_q"].contiguous().view(torch.uint8).reshape(B, 4)
    out[:, 8:12] = slot["reward"].contiguous().view(torch.uint8).reshape(B, 4)
    out[:, 12:24] = slot["visits"].contiguous().view(torch.uint8).reshape(B, 12)
    out[:, 24] = slot["action"]
    out[:, 25] = slot["flags"]
    return out


def unpack_records(buf: torch.Tensor) -> dict:
    n = buf.shape[0]
    return dict(state=buf[:, 0:4].contiguous().view(torch.int32).reshape(n),
                root_q=buf[:, 4:8].contiguous().view(torch.float32).reshape(n),
                reward=buf[:, 8:12].contiguous().view(torch.float32).reshape(n),
                visits=buf[:, 12:24].contiguous().view(torch.int16).reshape(n, 6),
                action=buf[:, 24].contiguous(), flags=buf[:, 25].contiguous())


def all_gather_records(slot: dict, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gathers one move's records from every rank -> uint8 [world * B, 26] (rank-major).
    Every rank must contribute the same B (weak scaling: fixed games per GPU)."""
    packed = pack_records(slot)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return packed
    world = dist.get_world_size()
    if out is None:
        out = torch.empty(world * packed.shape[0], 26, dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    return out
