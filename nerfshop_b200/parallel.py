"""Image-plane partition across GPUs (new design; the reference is single-GPU, README.md:423-425).

Rank r renders the 16x8-pixel tiles t with t % world == r (interleaved for load balance: contiguous bands put the
whole object on one GPU), packs them into a [n_tiles*128, 4] shard, ONE all-gather moves every shard to every
rank over NVLink, and each rank scatters the other ranks' shards into its framebuffer. No exchange inside the path.
"""
from __future__ import annotations

import numpy as np

TILE_W, TILE_H, TILE_PIXELS = 16, 8, 128


def n_tiles(W: int, H: int) -> int:
    return ((W + TILE_W - 1) // TILE_W) * ((H + TILE_H - 1) // TILE_H)


def tiles_for_rank(W: int, H: int, rank: int, world: int) -> int:
    return (n_tiles(W, H) + world - 1 - rank) // world


def max_tiles(W: int, H: int, world: int) -> int:
    return tiles_for_rank(W, H, 0, world)


def tile_pixel_indices(W: int, H: int, rank: int, world: int) -> np.ndarray:
    """Packed-shard position -> pixel index (row-major, -1 outside the image): the mirror of k_pack_tiles."""
    tiles_x = (W + TILE_W - 1) // TILE_W
    nt = tiles_for_rank(W, H, rank, world)
    local = np.arange(nt)[:, None]
    lane = np.arange(TILE_PIXELS)[None, :]
    tile = rank + local * world
    px = (tile % tiles_x) * TILE_W + lane % TILE_W
    py = (tile // tiles_x) * TILE_H + lane // TILE_W
    idx = np.where((px < W) & (py < H), px + W * py, -1)
    return idx.reshape(-1).astype(np.int64)


def unpack_numpy(shard: np.ndarray, W: int, H: int, rank: int, world: int, fb_flat: np.ndarray) -> None:
    idx = tile_pixel_indices(W, H, rank, world)
    valid = idx >= 0
    fb_flat[idx[valid]] = shard[: idx.size][valid]


def gather_framebuffer(renderer, fb, rank: int, world: int, shard, gathered, depth=None):
    """bench.py / a host's multi-GPU frame. RGBA and depth travel in ONE buffer: a shard row is 5 floats x 128 pixels per tile stored as
    [tiles*128 float4 | tiles*128 float] (shard: float32 [tiles_per_shard*128*5], gathered: [world, tiles_per_shard*128*5]); one pack launch,
    ONE all-gather, ONE unpack launch that scatters every other rank's tiles into fb and depth."""
    import torch.distributed as dist

    n = shard.numel() // 5  # pixels per shard
    rgba = shard[: 4 * n].view(n, 4)
    dpt = shard[4 * n:] if depth is not None else None
    renderer.pack_tiles(fb, depth, rank, world, rgba, dpt)
    dist.all_gather_into_tensor(gathered, shard.view(1, -1))
    g_rgba = gathered[:, : 4 * n]
    g_depth = gathered[:, 4 * n:] if depth is not None else None
    renderer.unpack_gathered(g_rgba, g_depth, gathered.stride(0), rank, world, fb, depth)
    return fb
