"""Multi-GPU plumbing of the frame-independent mode.

Frames shard across ranks with no data-path collective; the only exchange is one all-gather of the per-rank
keep-masks folded onto the global map.  Backend-agnostic (`nccl` on the GPUs, `gloo` in the CPU tests): these
helpers only move masks, the path itself runs in liberasor_b200.so.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous, balanced frame range of `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def fold_masks(n_global: int, voi_indices: Sequence[torch.Tensor], keep_masks: Sequence[torch.Tensor],
               device=None) -> torch.Tensor:
    """AND the per-frame keep masks (one byte per VoI point) onto the global map: a map point survives on this rank
    if no local frame rejected it.  voi_indices[f] maps frame f's VoI points to global map indices."""
    device = device if device is not None else (keep_masks[0].device if keep_masks else "cpu")
    keep_g = torch.ones(n_global, dtype=torch.uint8, device=device)
    if len(keep_masks):
        idx = torch.cat([i.to(device=device, dtype=torch.int64) for i in voi_indices])
        val = torch.cat([k.to(device=device, dtype=torch.uint8) for k in keep_masks])
        keep_g.scatter_reduce_(0, idx, val, reduce="amin")
    return keep_g


def allgather_and(keep_g: torch.Tensor, gather_buf: torch.Tensor | None = None) -> torch.Tensor:
    """The path's single collective: all-gather the per-rank masks and AND them (every rank gets the final mask).
    gather_buf: optional reusable (world, n) uint8 buffer (a streaming caller double-buffers it)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return keep_g
    world = dist.get_world_size()
    buf = gather_buf if gather_buf is not None else torch.empty((world, keep_g.numel()), dtype=torch.uint8, device=keep_g.device)
    dist.all_gather_into_tensor(buf.view(-1), keep_g.contiguous())
    return buf.amin(dim=0)


def static_map_mask(n_global: int, voi_indices: Sequence[np.ndarray], keep_masks: Sequence[np.ndarray], device="cpu") -> torch.Tensor:
    """fold + all-gather for numpy inputs (test convenience)."""
    vi = [torch.from_numpy(np.asarray(v).astype(np.int64)) for v in voi_indices]
    km = [torch.from_numpy(np.asarray(k).astype(np.uint8)) for k in keep_masks]
    return allgather_and(fold_masks(n_global, vi, km, device=device))
