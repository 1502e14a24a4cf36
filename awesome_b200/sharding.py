"""Sharding of independent fit units (frames, objects, sequences) over one-process-per-GPU ranks.

Per-frame prior fitting has no cross-unit dependency unless the reference's warm-start chain is on
(``reuse_state``: frame i starts from frame i-1, ``awesome/model/path_connected_net.py:867-870``).  Two
policies: ``"interleave"`` (unit u -> rank u % world, cold fits, best balance) and ``"chunk"`` (contiguous
chains per rank; the first frame of each chunk is cold, the rest may warm-start like the reference).
No collective runs on the data path; results are gathered once at the end."""
from __future__ import annotations

from typing import Any, List, Sequence


def shard_units(n_units: int, rank: int, world: int, policy: str = "interleave") -> List[int]:
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_units < 0:
        raise ValueError("n_units must be >= 0")
    if policy == "interleave":
        return list(range(rank, n_units, world))
    if policy == "chunk":
        base, rem = divmod(n_units, world)
        start = rank * base + min(rank, rem)
        return list(range(start, start + base + (1 if rank < rem else 0)))
    raise ValueError(f"unknown policy {policy!r}")


def gather_objects(local: Any, group=None) -> List[Any]:
    """All ranks' results on every rank (list indexed by rank); identity without torch.distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    out: List[Any] = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out


def merge_by_unit(per_rank: Sequence[dict]) -> dict:
    """Merge ``{unit_index: result}`` dicts coming from the ranks; a unit may only be owned once."""
    merged: dict = {}
    for d in per_rank:
        for k, v in (d or {}).items():
            if k in merged:
                raise ValueError(f"unit {k} was fitted by more than one rank")
            merged[k] = v
    return dict(sorted(merged.items()))
