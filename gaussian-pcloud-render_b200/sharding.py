"""Tile-row sharding of one frame across the ranks of a node.

Each rank holds the whole cloud, preprocesses every Gaussian, but bins and blends only a contiguous range of tile
rows; the ranges are chosen by prefix-sum balancing of a per-row cost so that the busiest rank is close to 1/N of
the frame (an equal split of tile rows leaves the busiest of 8 ranks with 22 % of the instances).
After blending, the ranks exchange their slabs so that every rank ends up with the full (3,H,W) image.

This module is host logic only (numpy + torch.distributed); the CUDA side is the `tile_rows` argument of the
rasterizer.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

TILE = 16


def views_of_rank(rank: int, world: int, num_views: int) -> List[int]:
    """View-parallel sharding: frames of an orbit are independent, rank r renders views
    r, r + world, r + 2*world, ...; no data-path collective is needed."""
    return list(range(rank, num_views, world))


def balanced_rows(row_cost: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous tile-row ranges [a,b) with ~equal summed cost.  Deterministic, so every rank can derive the same
    partition locally from the same cost vector; empty ranges are possible when rows are few or cost is zero."""
    n = len(row_cost)
    c = np.cumsum(np.asarray(row_cost, dtype=np.float64))
    total = c[-1] if n and c[-1] > 0 else 1.0
    cuts = [0]
    for k in range(1, world):
        cuts.append(int(np.searchsorted(c, total * k / world, side="left")) + 1)
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(np.asarray(cuts), n))
    cuts[-1] = n
    return [(int(cuts[k]), int(cuts[k + 1])) for k in range(world)]


def rebalance_rows(rows: Sequence[Tuple[int, int]], times_ms: Sequence[float], n_rows: int) -> List[Tuple[int, int]]:
    """Profile-guided refinement of a tile-row partition: given the time every rank took for ITS range of the same view,
    moves the range boundaries so that every rank gets a share of the rows inversely proportional to its time per row
    (what the cost model does not see -- the Gaussians a rank must still process because they reach into its rows from
    the neighbouring ranges, the fixed cost of the small sort / list kernels -- is in the measured times).
    Deterministic in its inputs, so every rank derives the same partition from the all-gathered times; ranges stay
    contiguous, ordered and cover [0, n_rows); a rank that had no rows keeps none."""
    world = len(rows)
    cnt = np.array([max(0, b - a) for a, b in rows], dtype=np.float64)
    t = np.maximum(np.asarray(times_ms, dtype=np.float64), 1e-6)
    speed = np.where(cnt > 0, cnt / t, 0.0)              # rows per ms of each rank on its own stretch of the frame
    if not np.any(speed > 0):
        return [tuple(r) for r in rows]
    # target: equal time T with rows_k = speed_k * T  ->  rows_k proportional to speed_k; move at most a quarter of a range
    want = speed / speed.sum() * cnt.sum()
    want = np.clip(want, 0.75 * cnt, 1.25 * cnt + 1.0)
    want *= cnt.sum() / max(want.sum(), 1e-9)
    cuts = np.concatenate([[0.0], np.cumsum(want)])
    cuts = np.rint(cuts).astype(np.int64)
    cuts[0], cuts[-1] = 0, int(n_rows)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n_rows))
    return [(int(cuts[k]), int(cuts[k + 1])) for k in range(world)]


def row_cost(need_tiles: np.ndarray, inst_tiles: np.ndarray) -> np.ndarray:
    """Per tile-row cost model: blended list prefix (need_t) + a share of the binned instances + a constant."""
    return need_tiles.sum(1).astype(np.float64) + 0.25 * inst_tiles.sum(1).astype(np.float64) + 8.0


def pixel_rows(rows: Sequence[Tuple[int, int]], H: int) -> List[Tuple[int, int]]:
    return [(min(H, a * TILE), min(H, b * TILE)) for a, b in rows]


def exchange_image(color: torch.Tensor, rows: Sequence[Tuple[int, int]], rank: int, group=None) -> torch.Tensor:
    """In-place assembly of the full image on every rank with ONE all-gather.

    `color` is (3,H,W); this rank rendered pixel rows pixel_rows(rows)[rank].  Row ranges are work-balanced and
    therefore unequal, so every rank contributes a slab padded to the tallest range; after the collective each
    rank copies the other ranks' slabs into place (plain device copies).  Bit-exact: pixels are only moved."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return color
    world = dist.get_world_size(group)
    H, W = int(color.shape[1]), int(color.shape[2])
    prow = pixel_rows(rows, H)
    tallest = max(b - a for a, b in prow)
    if tallest == 0:
        return color
    send = torch.zeros((3, tallest, W), dtype=color.dtype, device=color.device)
    a, b = prow[rank]
    if b > a:
        send[:, : b - a, :] = color[:, a:b, :]
    slabs = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(slabs, send, group=group)
    for r, (a, b) in enumerate(prow):
        if r != rank and b > a:
            color[:, a:b, :] = slabs[r][:, : b - a, :]
    return color


class PeerFrame:
    """Frame images in symmetric memory (torch.distributed._symmetric_memory): every rank can address every rank's
    (3,H,W) image through NVLink peer mappings, so the blend kernel writes its tile rows straight into all of them
    (GsScene.peer_out_color) and one symmetric-memory barrier replaces the image collective.  Images are
    double-buffered: a rank overwrites buffer k of its peers only two frames later, i.e. after a barrier that the
    peers reach once they are done with frame k (stream order), so one barrier per frame suffices."""

    def __init__(self, H: int, W: int, device, group=None, buffers: int = 2):
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.images, self.handles, self.peer_ptrs = [], [], []
        for _ in range(buffers):
            t = symm.empty((3, H, W), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, group)
            self.images.append(t)
            self.handles.append(h)
            self.peer_ptrs.append([h.get_buffer(r, (3, H, W), torch.float32).data_ptr() for r in range(self.world)])
        self.turn = 0

    def next(self):
        """(image of this rank, device pointers of all ranks' images, barrier callable) for the next frame."""
        k = self.turn % len(self.images)
        self.turn += 1
        return self.images[k], self.peer_ptrs[k], self.handles[k].barrier
