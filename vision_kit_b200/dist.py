"""Multi-GPU plumbing of the path: images shard across ranks with no collective on the hot
path; the only exchange is the all-gather of per-image detections for mAP eval (SURVEY.md
§8e).  One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_images` for `rank` (remainder to the first ranks)."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_images: int, world: int) -> List[int]:
    return [shard_range(n_images, r, world)[1] - shard_range(n_images, r, world)[0] for r in range(world)]


def allgather_detections(dets: torch.Tensor, counts: torch.Tensor, n_images: int = None, group=None):
    """dets (B_local, max_det, 6) float32 and counts (B_local,) int32 of this rank's shard ->
    (dets (n_images, max_det, 6), counts (n_images,)) in global image order on every rank.
    Shards may differ by one image (shard_range); they are padded to the largest for the
    collective.  Enqueued on the current stream: no host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()):
        return dets, counts
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if n_images is None:
        n_images = int(dets.shape[0]) * world
    sizes = shard_sizes(n_images, world)
    bmax = max(sizes)
    if int(dets.shape[0]) != sizes[rank]:
        raise ValueError(f"rank {rank} holds {int(dets.shape[0])} images, shard_range gives {sizes[rank]}")
    md = int(dets.shape[1])
    if min(sizes) == bmax and dets.is_contiguous() and counts.is_contiguous():
        # equal shards (the BASELINE configs): the tensors go out as they are, nothing is packed or copied first
        all_d = torch.empty((world * bmax, md, 6), dtype=dets.dtype, device=dets.device)
        all_c = torch.empty((world * bmax,), dtype=counts.dtype, device=counts.device)
        dist.all_gather_into_tensor(all_d, dets, group=group)
        dist.all_gather_into_tensor(all_c, counts, group=group)
        return all_d, all_c
    # one fused buffer per rank: detections + counts (as float bits) -> a single collective
    pack = torch.zeros((bmax, md * 6 + 1), dtype=torch.float32, device=dets.device)
    pack[: sizes[rank], : md * 6] = dets.reshape(sizes[rank], md * 6)
    pack[: sizes[rank], md * 6] = counts.to(torch.float32)
    flat = torch.empty((world * bmax, md * 6 + 1), dtype=torch.float32, device=dets.device)
    dist.all_gather_into_tensor(flat, pack, group=group)
    out = flat.view(world, bmax, md * 6 + 1)
    parts_d, parts_c = [], []
    for r in range(world):
        parts_d.append(out[r, : sizes[r], : md * 6].reshape(sizes[r], md, 6))
        parts_c.append(out[r, : sizes[r], md * 6].to(torch.int32))
    return torch.cat(parts_d, 0), torch.cat(parts_c, 0)
