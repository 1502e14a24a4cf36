"""Per-frame prior fitting of a whole sequence, sharded over one-process-per-GPU ranks (BASELINE configs[1]:
60 frames of 640x480 over 8 B200): the reference's loop ``PathConnectedNet._prior_based_pretrain``
(``awesome/model/path_connected_net.py:801-1007``) with its warm-start chain cut into a FIXED number of segments.

The reference chains every frame to its predecessor (``reuse_state``, ``:867-870``): one cold fit (``num_epochs``) and
T - 1 warm ones (``reuse_state_epochs``), strictly serial.  Sharding needs independent units, so the sequence is cut into
``n_segments`` contiguous segments (8 by default: one per GPU of a box) and the chain restarts -- cold, from a seeded
fresh prior -- at every segment start.  The segmentation does NOT depend on the number of ranks: segment s is fitted by
rank ``s % world`` with exactly the same arithmetic whatever ``world`` is, so the per-frame results (state, mask, IoU) of
a run on 1, 2, 4 or 8 GPUs are bit-identical and the speed-up is the pure distribution of equal work.  Inside a segment
the frames are chained one by one like in the reference (``group=1``, the default: every frame warm-starts from its
predecessor's last proper state); ``group=G`` fits G frames per fused launch from the group's entry state
(``fit_frames_grouped``), which makes the launches 15 % cheaper per frame but spends the 4 000 cold steps on G frames at
a time -- measured on the 60-frame sequence of configs[1]: 10.7 frames/s for G = 1, 7.7 for G = 2, 4.2 for G = 4 on one
B200, same mean IoU.  No-foreground skip, IoU check and per-frame retry as in the reference.

No collective runs during fitting; one ``all_gather_object`` of the per-frame results (bit-packed masks, fitted states,
IoU) at the end."""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import torch

from .pretrain import FitSchedule, fit_frames_grouped
from .sharding import gather_objects, merge_by_unit


def plan_segments(n_frames: int, n_segments: int) -> List[List[int]]:
    """Contiguous segments whose sizes differ by at most one (60 frames / 8 -> 8, 8, 8, 8, 7, 7, 7, 7)."""
    if n_frames < 0 or n_segments < 1:
        raise ValueError("n_frames must be >= 0 and n_segments >= 1")
    n_segments = min(n_segments, max(1, n_frames))
    base, rem = divmod(n_frames, n_segments)
    out, start = [], 0
    for s in range(n_segments):
        n = base + (1 if s < rem else 0)
        out.append(list(range(start, start + n)))
        start += n
    return out


def segments_of_rank(n_segments: int, rank: int, world: int) -> List[int]:
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_segments, world))


def _pack(mask: torch.Tensor) -> torch.Tensor:
    """bool [N] (device) -> uint8 [ceil(N/8)] on the host, MSB first (``synth.unpack_mask`` / ``numpy.unpackbits``)."""
    m = mask.reshape(-1).to(torch.uint8)
    pad = (-m.numel()) % 8
    if pad:
        m = torch.cat([m, m.new_zeros(pad)])
    w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=m.device)
    return (m.reshape(-1, 8) * w).sum(dim=1, dtype=torch.int32).to(torch.uint8).cpu()


def fit_sequence_sharded(prior_type, prior_args: Dict[str, Any], grid, unaries: Union[Sequence[torch.Tensor], Callable[[int], torch.Tensor]],
                         n_frames: int, schedule: Optional[FitSchedule] = None, n_segments: int = 8, group: int = 1,
                         rank: Optional[int] = None, world: Optional[int] = None, device=None, seed: int = 42,
                         gather: bool = True, keep_states: bool = True) -> Dict[int, Dict[str, Any]]:
    """Fit one prior per frame over the whole sequence; returns ``{frame: {"iou", "proper_fit", "skipped", "retries",
    "steps", "final_loss", "mask_fg_packed", "state", "segment", "rank"}}`` for ALL frames on every rank (``gather``) or
    for this rank's frames only.

    ``unaries``: the frames' soft segmentations (fg = 0 convention), a sequence or ``index -> tensor`` (host or device;
    only this rank's frames are touched).  ``prior_type(**prior_args)`` builds a prior WITHOUT a flow (grouped fits);
    segment s starts from the prior constructed under ``torch.manual_seed(seed + s)``."""
    import torch.distributed as dist
    from .model import NumberBasedMultiPriorModule
    s = schedule or FitSchedule()
    if rank is None or world is None:
        on = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    get = unaries if callable(unaries) else (lambda i: unaries[i])
    segs = plan_segments(n_frames, n_segments)
    multi = None
    local: Dict[int, Dict[str, Any]] = {}
    for si in segments_of_rank(len(segs), rank, world):
        frames = segs[si]
        if not frames:
            continue
        torch.manual_seed(seed + si)
        fresh = prior_type(**prior_args)                         # seeded on the host: the same cold start on any rank
        if multi is None:
            multi = NumberBasedMultiPriorModule(prior=fresh, min_priors=group).to(device)
        with torch.no_grad():
            for p in multi.priors:
                p.load_state_dict(fresh.state_dict())
        res = fit_frames_grouped(multi, grid, [get(i) for i in frames], s, frame_indices=frames, keep_masks=True)
        for r in res:
            local[r.index] = {
                "iou": r.iou, "proper_fit": r.proper_fit, "skipped": r.skipped, "retries": r.retries, "steps": r.steps,
                "final_loss": r.final_loss, "segment": si, "rank": rank,
                "mask_fg_packed": _pack(r.mask_fg) if r.mask_fg is not None else None,
                "state": r.state.detach().cpu() if (keep_states and r.state is not None) else None}
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    if not gather:
        return dict(sorted(local.items()))
    return merge_by_unit(gather_objects(local))
