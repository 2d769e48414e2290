"""World-size-2 gloo test (CPU) of the host-side multi-GPU logic: tile ownership t % world == rank, packed shard
sizes, and the single all-gather + unpack that assembles the framebuffer (bench.py's N > 1 path, with the CUDA
pack/unpack kernels replaced by their numpy index mirror nerfshop_b200.parallel.tile_pixel_indices)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerfshop_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = parallel.tile_pixel_indices(W, H, rank, world)          # [n_tiles*128] pixel index or -1
    n_max = parallel.max_tiles(W, H, world) * 128
    shard = torch.zeros((n_max, 4), dtype=torch.float32)
    valid = idx >= 0
    pix = torch.from_numpy(idx[valid].astype(np.int64))
    # "render": every owned pixel gets a value that identifies it
    vals = torch.stack([pix.float(), (pix % W).float(), (pix // W).float(), torch.full_like(pix, rank + 1).float()], 1)
    shard[: idx.size][torch.from_numpy(valid)] = vals
    gathered = [torch.zeros_like(shard) for _ in range(world)]
    dist.all_gather(gathered, shard)                                 # the ONE collective of the path
    fb = torch.zeros((H * W, 4))
    for k in range(world):
        parallel.unpack_numpy(gathered[k].numpy(), W, H, k, world, fb.numpy())
    if rank == 0:
        np.save(out_path, fb.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_assembles_the_frame(tmp_path):
    W, H, world = 203, 117, 2   # not a multiple of the 16x8 tile
    out = str(tmp_path / "fb.npy")
    mp.spawn(_worker, args=(world, _free_port(), W, H, out), nprocs=world, join=True)
    fb = np.load(out)
    pix = np.arange(W * H)
    assert np.array_equal(fb[:, 0], pix) and np.array_equal(fb[:, 1], pix % W) and np.array_equal(fb[:, 2], pix // W)
    tiles_x = (W + 15) // 16
    owner = (((pix // W) // 8) * tiles_x + (pix % W) // 16) % world
    assert np.array_equal(fb[:, 3], owner + 1)                     # every pixel written exactly by its owner


def test_partition_covers_every_pixel_once():
    for (W, H, world) in ((1920, 1080, 8), (203, 117, 3), (16, 8, 4), (17, 9, 2)):
        seen = np.zeros(W * H, np.int32)
        total_tiles = 0
        for r in range(world):
            idx = parallel.tile_pixel_indices(W, H, r, world)
            total_tiles += idx.size // 128
            seen[idx[idx >= 0]] += 1
        assert (seen == 1).all()
        assert total_tiles == ((W + 15) // 16) * ((H + 7) // 8)
        assert parallel.max_tiles(W, H, world) == max(parallel.tiles_for_rank(W, H, r, world) for r in range(world))
