"""Frame-stream plumbing: stereo pairs are independent, so a stream is sharded frame-by-frame over the GPUs of a
box (frame i -> rank i mod N) with NO collective on the data path; inside one GPU, frames
are pipelined over the engine's lanes.  torch.distributed is used only to agree on the timing (max over ranks)
and to gather per-frame results for re-ordering."""
from __future__ import annotations

from typing import Callable, Iterable, List, Sequence, Tuple


def shard(frame_ids: Sequence[int], rank: int, world: int) -> List[int]:
    """Round-robin shard of a frame stream: the frames rank `rank` of `world` processes."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(frame_ids[rank::world])


def owner(frame_id: int, world: int) -> int:
    return frame_id % world


def merge_in_order(per_rank: Sequence[Sequence[Tuple[int, object]]]) -> List[Tuple[int, object]]:
    """Re-orders (frame_id, result) pairs gathered from all ranks back into stream order."""
    out = [item for part in per_rank for item in part]
    out.sort(key=lambda t: t[0])
    ids = [t[0] for t in out]
    if len(set(ids)) != len(ids):
        raise ValueError("a frame was processed by more than one rank")
    return out


def run_lanes(n_lanes: int, frames: Iterable, enqueue: Callable[[int, object], None], wait: Callable[[int], None]) -> int:
    """Pipelines `frames` over `n_lanes` lanes: lane k holds frame k, k+n, ...; a lane is waited on right before it
    is reused and once more at the end.  Returns the number of frames issued."""
    n = 0
    for i, fr in enumerate(frames):
        ln = i % n_lanes
        if i >= n_lanes:
            wait(ln)
        enqueue(ln, fr)
        n += 1
    for ln in range(min(n_lanes, n)):
        wait(ln)
    return n


def reduce_throughput(n_frames_local: int, elapsed_s_local: float, dist=None):
    """Whole-job frames/s = all frames of all ranks / the slowest rank's time.  `dist` = torch.distributed (or None)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return n_frames_local / elapsed_s_local, n_frames_local, elapsed_s_local
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([elapsed_s_local], dtype=torch.float64, device=dev)
    n = torch.tensor([n_frames_local], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()) / float(t.item()), int(n.item()), float(t.item())
