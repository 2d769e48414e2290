"""Per-phase device times of the N > 1 frame (torchrun): render | pack | all-gather | unpack, CUDA events on the step's stream, max over ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from nerfshop_b200 import parallel, synthetic as syn
from nerfshop_b200.renderer import NerfRenderer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W, H = 1920, 1080
model = syn.make_model(seed=1337)
r = NerfRenderer(local)
r.upload_model(model.desc, model.params)
r.upload_occupancy(syn.make_occupancy(model))
cams = syn.orbit_cameras(120)
fb = torch.zeros((H, W, 4), device=dev)
depth = torch.zeros((H, W), device=dev)
n = r.tiles_for_rank(W, H, 0, world) * 128
shard = torch.zeros(5 * n, device=dev)
gathered = torch.zeros((world, 5 * n), device=dev)
names = ["zero", "render", "pack", "all_gather", "unpack"]
acc = np.zeros(len(names))
for it in range(14):
    f = syn.make_frame(model, cams[(it * 7) % 120], W, H, rank=rank, world=world)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    fb.zero_()
    ev[1].record()
    r.render(f, fb, depth)
    ev[2].record()
    r.pack_tiles(fb, depth, rank, world, shard[: 4 * n].view(n, 4), shard[4 * n:])
    ev[3].record()
    dist.all_gather_into_tensor(gathered, shard.view(1, -1))
    ev[4].record()
    r.unpack_gathered(gathered[:, : 4 * n], gathered[:, 4 * n:], gathered.stride(0), rank, world, fb, depth)
    ev[5].record()
    torch.cuda.synchronize()
    if it >= 4:
        acc += np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(len(names))])
t = torch.tensor(acc / 10, dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: " + "  ".join(f"{nm} {v:.3f} ms" for nm, v in zip(names, t.tolist())) + f"  | sum {t.sum().item():.3f} ms; shard {5 * n * 4 / 1e6:.1f} MB per rank")
    print("NCCL env:", {k: v for k, v in os.environ.items() if k.startswith("NCCL")})
dist.destroy_process_group()
