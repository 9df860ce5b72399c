"""Tile partitioning of one frame across processes (one process per GPU) and the gather of the pieces.

This is the in-box replacement of the reference's TCP master/worker pair (reference
VerStarting/main_net_master.cc:195-236, main_net_worker.cc:127-166): the frame is cut into tiles, every
worker holds the whole scene and renders the tiles it owns, the master blits them into the frame
(BlitWorkChunk, main_net_master.cc:223-236).  Here a tile is a strip of 8 image rows, strips are dealt
round-robin (strip s belongs to rank s % world -- the same rule libmythtracer_b200 applies to the devices
of one context, see mtb_set_partition), and the blit is one gather over NCCL / NVLink.

torch.distributed is plumbing only: the pixels are produced by the CUDA kernels of libmythtracer_b200.so.
"""
from __future__ import annotations

STRIP_ROWS = 8


def n_strips(height: int) -> int:
    return (height + STRIP_ROWS - 1) // STRIP_ROWS


def strip_owner(strip: int, world: int) -> int:
    return strip % world


def owned_strips(height: int, rank: int, world: int):
    """Strip indices rank `rank` of `world` renders."""
    return list(range(rank, n_strips(height), world))


def owned_rows(height: int, rank: int, world: int):
    rows = []
    for s in owned_strips(height, rank, world):
        rows.extend(range(s * STRIP_ROWS, min(height, (s + 1) * STRIP_ROWS)))
    return rows


def padded_height(height: int, world: int) -> int:
    """Rows of the per-rank frame buffer: a whole number of strip rounds, so that the buffer can be viewed
    as [rounds, world, STRIP_ROWS * width * 3]."""
    rounds = (n_strips(height) + world - 1) // world
    return rounds * world * STRIP_ROWS


def gather_frame(local, height: int, width: int, rank: int, world: int, dst: int = 0, group=None, out=None):
    """Assembles the frame on rank `dst`.

    local: uint8 tensor [padded_height(height, world), width, 3] in which this rank's strips are rendered
    (other rows are ignored).  Returns a uint8 tensor [height, width, 3] on `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    hp = padded_height(height, world)
    assert local.dtype == torch.uint8 and tuple(local.shape) == (hp, width, 3), (local.shape, hp, width)
    if world == 1:
        return local[:height]
    strip_bytes = STRIP_ROWS * width * 3
    rounds = hp // (world * STRIP_ROWS)
    mine = local.view(rounds, world, strip_bytes)[:, rank].contiguous()
    if rank == dst:
        pieces = [torch.empty_like(mine) for _ in range(world)]
        dist.gather(mine, pieces, dst=dst, group=group)
        frame = out if out is not None else torch.empty((hp, width, 3), dtype=torch.uint8, device=local.device)
        view = frame.view(rounds, world, strip_bytes)
        for r in range(world):
            view[:, r].copy_(pieces[r])
        return frame[:height]
    dist.gather(mine, None, dst=dst, group=group)
    return None
