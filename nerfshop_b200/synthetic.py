"""Seeded synthetic inputs for the render path (no checkpoint ships with the reference).

Everything here is plain numpy and deterministic in `seed` (reference default seed 1337, testbed.h:507):
  * a base.json-shaped parameter block [density MLP | rgb MLP | hash grid] whose density field is a
    fox-sized blob scene (body/head/legs/tail + ground slab + partial background dome), carried by the
    dense level 3 of the hash grid and a hand-set path through the density MLP; all other weights and
    hash features are seeded noise so every level and every MLP row is exercised;
  * the matching 5x128^3 occupancy bitfield (Morton order per cascade, common_nerf.cu:117-141);
  * cameras: fox frame 0 through nerf_matrix_to_ngp (nerf_loader.h:74-92) and a 120-view orbit.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import abi

FOX_AABB_SCALE = 4
N_MLP_PARAMS = 64 * 32 + 16 * 64 + 64 * 32 + 64 * 64 + 16 * 64  # 10240

# data/nerf/fox/transforms.json frame 0 ("images/0001.jpg") transform_matrix, rows 0-2
FOX_FRAME0 = np.array(
    [
        [0.8926439112348871, 0.08799600283226543, 0.4420900262071262, 3.168359405609479],
        [0.4464189982715247, -0.03675452191179031, -0.8940689141475064, -5.4794898611466945],
        [-0.062425682580756266, 0.995442519072023, -0.07209178487538156, -0.9791660699008925],
    ],
    dtype=np.float64,
)


def per_level_scale(aabb_scale: int, n_levels: int = 16, base_resolution: int = 16) -> float:
    """testbed.cu:2280-2292: exp(log(2048*aabb_scale/N_min)/(L-1)), evaluated in float like the reference."""
    v = np.float32(2048.0) * np.float32(aabb_scale) / np.float32(base_resolution)
    return float(np.exp(np.log(v, dtype=np.float32) / np.float32(n_levels - 1), dtype=np.float32))


def model_desc(aabb_scale: int = FOX_AABB_SCALE) -> abi.NsbModelDesc:
    return abi.NsbModelDesc(16, 2, 19, 16, per_level_scale(aabb_scale), 64, 1, 2, 4)


def level_table(desc: abi.NsbModelDesc):
    """tcnn GridEncoding level geometry. Returns scale f32[L], res, offset, size (entries)."""
    L = desc.n_levels
    l2s = np.log2(np.float32(desc.per_level_scale), dtype=np.float32)
    scale = np.zeros(L, np.float32)
    res = np.zeros(L, np.uint32)
    offset = np.zeros(L, np.uint32)
    size = np.zeros(L, np.uint32)
    off = 0
    for l in range(L):
        s = np.float32(np.exp2(np.float32(l) * l2s, dtype=np.float32) * np.float32(desc.base_resolution) - np.float32(1.0))
        r = int(math.ceil(float(s))) + 1
        n = min(((r ** 3 + 7) // 8) * 8, 1 << desc.log2_hashmap_size)
        scale[l], res[l], offset[l], size[l] = s, r, off, n
        off += n
    return scale, res, offset, size, off


# ---------------------------------------------------------------------------------------------
# scene geometry (NGP world coordinates, y up, object around (0.5,0.5,0.5))
# ---------------------------------------------------------------------------------------------
def _sd_ellipsoid(p, c, r):
    q = (p - np.asarray(c, np.float32)) / np.asarray(r, np.float32)
    k = np.linalg.norm(q, axis=-1)
    return (k - 1.0) * float(min(r))


def _sd_box(p, c, h):
    q = np.abs(p - np.asarray(c, np.float32)) - np.asarray(h, np.float32)
    return np.linalg.norm(np.maximum(q, 0.0), axis=-1) + np.minimum(np.max(q, axis=-1), 0.0)


def scene_inside(p: np.ndarray) -> np.ndarray:
    """v(x) in [-1,1]: +1 deep inside matter, -1 in free space (a clamped, banded signed distance)."""
    p = p.astype(np.float32)
    fox = np.minimum.reduce(
        [
            _sd_ellipsoid(p, (0.5, 0.47, 0.5), (0.20, 0.15, 0.30)),   # body
            _sd_ellipsoid(p, (0.5, 0.62, 0.78), (0.12, 0.12, 0.12)),  # head
            _sd_ellipsoid(p, (0.44, 0.76, 0.80), (0.04, 0.07, 0.04)),  # ears
            _sd_ellipsoid(p, (0.56, 0.76, 0.80), (0.04, 0.07, 0.04)),
            _sd_ellipsoid(p, (0.5, 0.55, 0.12), (0.06, 0.06, 0.18)),  # tail
            _sd_ellipsoid(p, (0.41, 0.33, 0.33), (0.045, 0.13, 0.045)),  # legs
            _sd_ellipsoid(p, (0.59, 0.33, 0.33), (0.045, 0.13, 0.045)),
            _sd_ellipsoid(p, (0.41, 0.33, 0.67), (0.045, 0.13, 0.045)),
            _sd_ellipsoid(p, (0.59, 0.33, 0.67), (0.045, 0.13, 0.045)),
        ]
    )
    ground = _sd_box(p, (0.5, -0.05, 0.5), (1.2, 0.07, 1.2))  # top at y = 0.02, below every camera
    # background dome: shell of radius 1.8 about (0.5,0.2,0.5), open around the azimuth of fox camera 0
    # (which sits outside it) so that every camera looks at the dome's far, inner side
    q = p - np.asarray((0.5, 0.2, 0.5), np.float32)
    rad = np.linalg.norm(q, axis=-1)
    dome = np.abs(rad - 1.8) - 0.15
    az = np.arctan2(q[..., 2], q[..., 0])
    daz = np.abs(np.angle(np.exp(1j * (az - 2.62))))
    dome = np.where(daz < 1.0, np.maximum(dome, 0.25), dome)
    v_near = np.clip(-np.minimum(fox, ground) / 0.06, -1.0, 1.0)
    v_far = np.clip(-dome / 0.10, -1.0, 1.0)
    return np.maximum(v_near, v_far).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# model parameters
# ---------------------------------------------------------------------------------------------
CARRIER_LEVEL = 3  # dense level (res 56 for fox) that carries v(x) in feature 0 and the constant 1 in feature 1


@dataclass
class SyntheticModel:
    desc: abi.NsbModelDesc
    params: np.ndarray  # uint16 view of fp16, reference block order
    aabb_min: np.ndarray
    aabb_max: np.ndarray
    aabb_scale: int

    @property
    def grid(self) -> np.ndarray:
        return self.params[N_MLP_PARAMS:].view(np.float16)


def make_model(seed: int = 1337, aabb_scale: int = FOX_AABB_SCALE) -> SyntheticModel:
    desc = model_desc(aabb_scale)
    scale, res, offset, size, total = level_table(desc)
    rng = np.random.Generator(np.random.Philox(seed))
    half = 0.5 * min(1 << (abi.NSB_NERF_CASCADES - 1), aabb_scale)
    aabb_min = np.full(3, 0.5 - half, np.float32)  # testbed_nerf.cu:3410-3411
    aabb_max = np.full(3, 0.5 + half, np.float32)

    # ---- hash grid: seeded noise everywhere, amplitude decaying with level ----
    grid = np.empty(total * 2, np.float16)
    for l in range(desc.n_levels):
        amp = 0.5 * (0.85 ** l)
        n = int(size[l]) * 2
        grid[int(offset[l]) * 2 : int(offset[l]) * 2 + n] = rng.uniform(-amp, amp, n).astype(np.float16)
    # ---- carrier level: feature 0 = v(x) at the level's vertices, feature 1 = 1 ----
    l = CARRIER_LEVEL
    r = int(res[l])
    assert r ** 3 <= int(size[l]), "carrier level must be dense"
    idx = np.arange(r, dtype=np.float32)
    xw = (idx - 0.5) / scale[l]  # vertex i sits at warped coordinate (i - 0.5)/scale  (pos = x*scale + 0.5)
    world = aabb_min[0] + xw * (aabb_max[0] - aabb_min[0])
    Z, Y, X = np.meshgrid(world, world, world, indexing="ij")
    v = scene_inside(np.stack([X, Y, Z], axis=-1))  # [z][y][x] -> dense index x + y*r + z*r*r
    lvl = grid[int(offset[l]) * 2 : int(offset[l]) * 2 + int(size[l]) * 2].reshape(-1, 2)
    lvl[: r ** 3, 0] = v.reshape(-1).astype(np.float16)
    lvl[:, 1] = np.float16(1.0)

    # ---- MLPs: Xavier-uniform noise + a hand-set density path  raw_sigma = 7 v - 1 ----
    def xavier(n_out, n_in, gain=1.0):
        a = gain * math.sqrt(6.0 / (n_in + n_out))
        return rng.uniform(-a, a, (n_out, n_in)).astype(np.float32)

    k_v, k_one = CARRIER_LEVEL * 2, CARRIER_LEVEL * 2 + 1
    W1 = xavier(64, 32, 1.0)
    W1[:, k_v] *= 0.25
    W1[:, k_one] *= 0.25
    W1[0, :] = 0.0
    W1[0, k_v], W1[0, k_one] = 7.0, 7.0  # h0 = relu(7 v + 7) = 7 v + 7
    W1[1, :] = 0.0
    W1[1, k_one] = 8.0                   # h1 = 8
    W2 = xavier(16, 64, 1.0)
    W2[0, :] *= 0.05
    W2[0, 0], W2[0, 1] = 1.0, -1.0       # out0 = h0 - h1 (+ small noise) = 7 v - 1
    W2[1:, 0:2] *= 0.1
    W3 = xavier(64, 32, 1.6)
    W3[:, 0] *= 0.1                      # keep the (large) raw density from saturating the colour net
    W4 = xavier(64, 64, 1.6)
    W5 = xavier(16, 64, 2.0)
    mlp = np.concatenate([w.reshape(-1) for w in (W1, W2, W3, W4, W5)]).astype(np.float16)
    assert mlp.size == N_MLP_PARAMS
    params = np.concatenate([mlp, grid]).view(np.uint16)
    return SyntheticModel(desc, np.ascontiguousarray(params), aabb_min, aabb_max, aabb_scale)


# ---------------------------------------------------------------------------------------------
# occupancy bitfield
# ---------------------------------------------------------------------------------------------
def _expand_bits(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3d(x, y, z):
    return _expand_bits(x) | (_expand_bits(y) << np.uint32(1)) | (_expand_bits(z) << np.uint32(2))


def _carrier_trilinear(model: SyntheticModel, world: np.ndarray) -> np.ndarray:
    """F(x): trilinear interpolation of the carrier level's feature 0 (what the density MLP sees)."""
    scale, res, offset, size, _ = level_table(model.desc)
    l = CARRIER_LEVEL
    r = int(res[l])
    table = model.grid[int(offset[l]) * 2 : int(offset[l]) * 2 + int(size[l]) * 2].reshape(-1, 2)[:, 0].astype(np.float32)
    xw = (world - model.aabb_min) / (model.aabb_max - model.aabb_min)
    p = xw.astype(np.float32) * scale[l] + np.float32(0.5)
    g = np.floor(p)
    w = p - g
    g = g.astype(np.int64)
    out = np.zeros(world.shape[:-1], np.float32)
    for c in range(8):
        wt = np.ones(world.shape[:-1], np.float32)
        gi = []
        for d in range(3):
            if c & (1 << d):
                wt = wt * w[..., d]
                gi.append(g[..., d] + 1)
            else:
                wt = wt * (1 - w[..., d])
                gi.append(g[..., d])
        index = (gi[0] + gi[1] * r + gi[2] * r * r) % int(size[l])
        out += wt * table[index]
    return out


def make_occupancy(model: SyntheticModel, v_threshold: float = -0.3) -> np.ndarray:
    """uint8[5*128^3/8]; cascade m covers the cube of side 2^m about 0.5 (common_nerf.cu:117-141)."""
    G = abi.NSB_NERF_GRIDSIZE
    max_cascade = 0
    while (1 << max_cascade) < model.aabb_scale:
        max_cascade += 1  # testbed_nerf.cu:3418-3421
    bits = np.zeros((abi.NSB_NERF_CASCADES, G ** 3), bool)
    ax = np.arange(G, dtype=np.float32)
    prev = None
    for m in range(max_cascade + 1):
        c = ((ax + 0.5) / G - 0.5) * (2.0 ** m) + 0.5
        Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
        F = _carrier_trilinear(model, np.stack([X, Y, Z], axis=-1))
        occ = F > v_threshold
        # dilate by one cell so the surface band is fully covered
        d = occ.copy()
        for axis in range(3):
            for sh in (1, -1):
                rolled = np.roll(occ, sh, axis=axis)
                sl = [slice(None)] * 3
                sl[axis] = 0 if sh == 1 else -1
                rolled[tuple(sl)] = False
                d |= rolled
        occ = d
        if prev is not None:  # bitfield_max_pool (testbed_nerf.cu:534-555): inner half = max-pool of the finer cascade
            pooled = prev.reshape(G // 2, 2, G // 2, 2, G // 2, 2).any(axis=(1, 3, 5))
            q = G // 4
            occ[q : q + G // 2, q : q + G // 2, q : q + G // 2] |= pooled
        prev = occ
        zi, yi, xi = np.nonzero(occ)
        bits[m, morton3d(xi, yi, zi)] = True
    packed = np.packbits(bits.reshape(-1), bitorder="little")
    assert packed.size == abi.NSB_BITFIELD_BYTES
    return packed


# ---------------------------------------------------------------------------------------------
# cameras and frames
# ---------------------------------------------------------------------------------------------
def nerf_matrix_to_ngp(m34: np.ndarray, scale: float = 0.33, offset=(0.5, 0.5, 0.5)) -> np.ndarray:
    """nerf_loader.h:74-92 (from_mitsuba = false)."""
    r = np.array(m34, dtype=np.float64).copy()
    r[:, 1] *= -1
    r[:, 2] *= -1
    r[:, 3] = r[:, 3] * scale + np.asarray(offset)
    r = r[[1, 2, 0], :]  # cycle axes xyz <- yzx
    return r.astype(np.float32)


def fox_camera0() -> np.ndarray:
    return nerf_matrix_to_ngp(FOX_FRAME0)


def look_at(pos, target=(0.5, 0.5, 0.5), up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """3x4 camera: columns = right (+u), down (+v), forward, position."""
    pos = np.asarray(pos, np.float64)
    f = np.asarray(target, np.float64) - pos
    f /= np.linalg.norm(f)
    upv = np.asarray(up, np.float64)
    down = -(upv - f * np.dot(upv, f))
    down /= np.linalg.norm(down)
    right = np.cross(down, f)
    return np.stack([right, down, f, pos], axis=1).astype(np.float32)


def orbit_cameras(n: int = 120, radius: float = 1.45, height: float = 0.35):
    cams = []
    for i in range(n):
        a = 2.0 * math.pi * i / n
        cams.append(look_at((0.5 + radius * math.cos(a), 0.5 + height, 0.5 + radius * math.sin(a))))
    return cams


def make_frame(model: SyntheticModel, camera: np.ndarray, width: int = 1920, height: int = 1080, spp: int = 0,
               rank: int = 0, world: int = 1, apply_operators: bool = False, render_mode: int = abi.NSB_RENDER_SHADE,
               poisson_target: bool = False) -> abi.NsbFrame:
    """Defaults of the reference's interactive render: focal = relative_focal_length(1) * res[fov_axis=1]
    (testbed.h:465-466, testbed.cu:2556-2558), screen centre 0.5, no rolling shutter, min transmittance 0.01
    (testbed.h:651), cone angle 1/256 for aabb_scale > 1 (testbed_nerf.cu:3425), Logistic rgb / Exponential
    density (:3341), depth_scale = 1/dataset.scale = 1/0.33."""
    f = abi.NsbFrame()
    f.width, f.height = width, height
    f.focal_length[0] = f.focal_length[1] = float(height)
    f.screen_center[0] = f.screen_center[1] = 0.5
    cam = np.asarray(camera, np.float32).reshape(3, 4)
    flat = cam.T.reshape(-1)  # column-major
    for i in range(12):
        f.camera0[i] = float(flat[i])
        f.camera1[i] = float(flat[i])
    for i in range(3):
        f.render_aabb_min[i] = float(model.aabb_min[i])
        f.render_aabb_max[i] = float(model.aabb_max[i])
        f.train_aabb_min[i] = float(model.aabb_min[i])
        f.train_aabb_max[i] = float(model.aabb_max[i])
    f.cone_angle_constant = 0.0 if model.aabb_scale <= 1 else 1.0 / 256.0
    f.min_transmittance = 0.01
    f.depth_scale = 1.0 / 0.33
    f.rgb_activation = abi.NSB_ACT_LOGISTIC
    f.density_activation = abi.NSB_ACT_EXPONENTIAL
    f.render_mode = render_mode
    f.spp_index = spp
    f.snap_to_pixel_centers = 0
    f.apply_operators = int(apply_operators)
    f.poisson_target = int(poisson_target)
    f.linear_colors = 0
    f.min_mip = 0
    f.tile_rank, f.tile_world = rank, world
    return f


# ---------------------------------------------------------------------------------------------
# ABI 3 frame extras: lens distortion, depth of field, glow, environment and distortion maps
# ---------------------------------------------------------------------------------------------
def make_envmap(width: int = 64, height: int = 32, seed: int = 11) -> np.ndarray:
    """A smooth latitude-longitude RGBA map [height, width, 4] (float32, alpha 1): what m_envmap holds (envmap.cuh:30-62)."""
    rng = np.random.default_rng(seed)
    u = np.linspace(0.0, 2.0 * np.pi, width, endpoint=False, dtype=np.float32)[None, :]
    v = np.linspace(0.0, np.pi, height, dtype=np.float32)[:, None]
    out = np.ones((height, width, 4), np.float32)
    for c in range(3):
        a, b, p = rng.uniform(0.1, 0.4, 3).astype(np.float32)
        out[..., c] = 0.5 + a * np.cos((c + 1) * u + p) * np.sin(v) + b * np.cos(2.0 * v + p)
    return np.ascontiguousarray(out)


def make_distortion_map(width: int = 32, height: int = 18, amplitude: float = 0.01, seed: int = 12) -> np.ndarray:
    """A smooth per-pixel offset of the camera-space ray direction [height, width, 2] (m_distortion.map, pixel_to_ray :278-280)."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, width, dtype=np.float32)[None, :]
    y = np.linspace(0.0, 1.0, height, dtype=np.float32)[:, None]
    p = rng.uniform(0.0, np.pi, 4).astype(np.float32)
    out = np.zeros((height, width, 2), np.float32)
    out[..., 0] = amplitude * np.sin(2.0 * np.pi * x + p[0]) * np.cos(np.pi * y + p[1])
    out[..., 1] = amplitude * np.cos(np.pi * x + p[2]) * np.sin(2.0 * np.pi * y + p[3])
    return np.ascontiguousarray(out)


def set_camera_distortion(f: abi.NsbFrame, mode: int, params) -> abi.NsbFrame:
    f.camera_distortion_mode = mode
    for i in range(7):
        f.camera_distortion_params[i] = float(params[i]) if i < len(params) else 0.0
    return f


def set_maps(f: abi.NsbFrame, envmap_ptr: int = 0, envmap_shape=None, distortion_ptr: int = 0, distortion_shape=None) -> abi.NsbFrame:
    """Pointers are what the consumer can read: device pointers for nsb_render, host pointers for a CPU consumer (the tests' checker)."""
    f.envmap_dev = envmap_ptr or None
    if envmap_shape is not None:
        f.envmap_resolution[0], f.envmap_resolution[1] = int(envmap_shape[1]), int(envmap_shape[0])
    f.distortion_dev = distortion_ptr or None
    if distortion_shape is not None:
        f.distortion_resolution[0], f.distortion_resolution[1] = int(distortion_shape[1]), int(distortion_shape[0])
    return f
