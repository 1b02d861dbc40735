"""Multi-GPU host logic: the index space [0, N) is cut into
contiguous equal shards, one per rank; every output element depends on one element of
each input, so there is NO collective on the data path.  ``torch.distributed`` is used
only for the start barrier, the max-over-ranks of the device times and the combination
of per-shard digests (sum / xor are order-independent).

The reference itself scales by replica pods only (cuda-test-hpa.yaml:11-12), each pod an
independent vectorAdd loop on its own GPU (cuda-test-deployment.yaml:20-22).
"""
from __future__ import annotations

import os

from .capi import shard_range  # noqa: F401  (re-export: the C-ABI arithmetic is the single source)

_MASK = (1 << 64) - 1


def world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process default)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str) -> bool:
    """Join the process group if launched under torchrun. Returns True if distributed."""
    import torch.distributed as dist

    rank, ws, _ = world()
    if ws <= 1:
        return False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    return True


def _reduce_device():
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def barrier() -> None:
    import torch
    import torch.distributed as dist

    if dist.is_initialized():
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=_reduce_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=_reduce_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def combine_digests(local: tuple[int, int]) -> tuple[int, int]:
    """Global (sum mod 2^64, xor) of the per-shard digests."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return local[0] & _MASK, local[1] & _MASK
    ws = dist.get_world_size()
    # 64-bit values travel as four 16-bit limbs so no backend's integer reduce can overflow
    limbs = [(local[0] >> s) & 0xFFFF for s in (0, 16, 32, 48)] + [(local[1] >> s) & 0xFFFF for s in (0, 16, 32, 48)]
    mine = torch.tensor(limbs, dtype=torch.int64, device=_reduce_device())
    every = [torch.zeros_like(mine) for _ in range(ws)]
    dist.all_gather(every, mine)
    s, x = 0, 0
    for t in every:
        v = [int(u) for u in t.tolist()]
        s = (s + (v[0] | v[1] << 16 | v[2] << 32 | v[3] << 48)) & _MASK
        x ^= v[4] | v[5] << 16 | v[6] << 32 | v[7] << 48
    return s, x


def gather_ints(values: list[int]) -> list[list[int]]:
    """Every rank's small list of ints, on every rank (diagnostics only)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return [list(values)]
    mine = torch.tensor(values, dtype=torch.int64, device=_reduce_device())
    every = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    return [[int(v) for v in t.tolist()] for t in every]
