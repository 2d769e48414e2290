"""Python host mirror of the reference surfaces the render path is called through, over the C ABI.

  NerfRenderer.render            <- Testbed::render_nerf            (testbed_nerf.cu:3066)
  NerfRenderer.render_to_cpu     <- Testbed::render_to_cpu          (python_api.cu:129-175)
  NerfRenderer.inference / density <- NerfNetwork::inference_mixed_precision / ::density
  NerfRenderer.map_rays / poisson_residuals <- EditOperator::map_rays / ::compute_poisson_full_residuals
  NerfRenderer.{add,reset}_edit_operator <- NerfTracer::{add,reset}_edit_operator (testbed.h:129-240)

torch is used for device memory and streams only; every computation is a call into libnerfshop_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


def _torch():
    import torch

    return torch


class NerfRenderer:
    def __init__(self, device: int | None = None, lib_path: str | None = None):
        """lib_path: an experiment build of the same library (nerfshop_b200.build.build_variant); default = the product library."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise abi.NsbError("no CUDA device: nerfshop_b200 has no CPU fallback")
        self.lib = abi.load_library(lib_path)
        self.device = torch.cuda.current_device() if device is None else device
        self.ctx = C.c_void_p()
        abi.check(self.lib, self.lib.nsb_create(self.device, C.byref(self.ctx)), "nsb_create")
        self._ops = []
        self._keep = None

    def close(self):
        if self.ctx:
            self.lib.nsb_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state -------------------------------------------------------------------------------------
    def upload_model(self, desc: abi.NsbModelDesc, params_u16: np.ndarray):
        p = np.ascontiguousarray(params_u16, dtype=np.uint16)
        abi.check(self.lib, self.lib.nsb_upload_model(self.ctx, C.byref(desc), p.ctypes.data, p.size), "nsb_upload_model")

    def set_mlp_accumulator(self, policy: int):
        """NSB_MLP_ACC_F16 (default: the arithmetic of tiny-cuda-nn's wmma __half accumulator fragments) or NSB_MLP_ACC_F32."""
        abi.check(self.lib, self.lib.nsb_set_mlp_accumulator(self.ctx, int(policy)), "nsb_set_mlp_accumulator")

    def upload_occupancy(self, bitfield: np.ndarray):
        b = np.ascontiguousarray(bitfield, dtype=np.uint8)
        abi.check(self.lib, self.lib.nsb_upload_occupancy(self.ctx, b.ctypes.data, b.size), "nsb_upload_occupancy")

    def upload_density_grid(self, density_grid: np.ndarray) -> np.ndarray:
        """Snapshot path: float[5*128^3] density grid -> occupancy bitfield (returned) — update_density_grid_mean_and_bitfield."""
        g = np.ascontiguousarray(density_grid, dtype=np.float32).reshape(-1)
        out = np.zeros(abi.NSB_BITFIELD_BYTES, np.uint8)
        abi.check(self.lib, self.lib.nsb_upload_density_grid(self.ctx, g.ctypes.data, g.size, out.ctypes.data), "nsb_upload_density_grid")
        return out

    def update_density_grid(self, rng, ema_step: int, n_uniform: int, n_nonuniform: int = 0, reset_grid: bool = False, n_cascades: int = 3,
                            decay: float = 0.95, train_aabb=((-1.5, -1.5, -1.5), (2.5, 2.5, 2.5)), density_activation: int = abi.NSB_ACT_EXPONENTIAL,
                            apply_operators: bool = True, stream: int = 0) -> None:
        """Testbed::update_density_grid_nerf_operator (testbed_nerf.cu:3533-3639): one update of the context's running density
        grid and occupancy bitfield through the uploaded operators. `rng` (nerfshop_b200.rng.Pcg32) is advanced like m_rng."""
        u = abi.NsbGridUpdate()
        u.n_uniform_samples, u.n_nonuniform_samples = int(n_uniform), int(n_nonuniform)
        u.reset_grid, u.n_cascades, u.decay, u.ema_step = int(bool(reset_grid)), int(n_cascades), float(decay), int(ema_step)
        u.rng_state, u.rng_inc = rng.state, rng.inc
        u.train_aabb_min[:] = train_aabb[0]
        u.train_aabb_max[:] = train_aabb[1]
        u.density_activation, u.apply_operators = int(density_activation), int(bool(apply_operators))
        abi.check(self.lib, self.lib.nsb_update_density_grid(self.ctx, C.byref(u), stream), "nsb_update_density_grid")
        rng.advance()  # :3576
        rng.advance()  # :3589
        return u

    def update_density_grid_nerf_render(self, state: dict, n_iterations: int, reset_grid: bool, **kw) -> None:
        """Testbed::update_density_grid_nerf_render (testbed_nerf.cu:3514-3520): full-grid uniform updates after an edit.
        `state` carries Testbed's members {"rng": Pcg32, "ema_step": int, "max_cascade": int}."""
        n_casc = state["max_cascade"] + 1
        for i in range(n_iterations):
            self.update_density_grid(state["rng"], state["ema_step"], 128 ** 3 * n_casc, 0, reset_grid and i == 0, n_cascades=n_casc, **kw)
            state["ema_step"] += 1  # :3636
        _torch().cuda.synchronize(self.device)

    def download_density_grid(self, want_grid: bool = True, want_bitfield: bool = True):
        grid = np.zeros(abi.NSB_GRID_CELLS, np.float32) if want_grid else None
        bits = np.zeros(abi.NSB_BITFIELD_BYTES, np.uint8) if want_bitfield else None
        abi.check(self.lib, self.lib.nsb_download_density_grid(self.ctx, None if grid is None else grid.ctypes.data, None if bits is None else bits.ctypes.data),
                  "nsb_download_density_grid")
        return grid, bits

    # ---- per-edit rebuild on the device (GrowingSelection::update_tet_mesh, growing_selection.cu:1615) ----
    def cage_attach_mvc(self, op_index: int, mvc_weights: np.ndarray) -> None:
        w = np.ascontiguousarray(mvc_weights, np.float32)
        abi.check(self.lib, self.lib.nsb_cage_attach_mvc(self.ctx, op_index, w.ctypes.data, w.shape[1]), "nsb_cage_attach_mvc")

    def cage_deform(self, op_index: int, cage_vertices: np.ndarray, stream: int = 0) -> None:
        cv = np.ascontiguousarray(cage_vertices, np.float32).reshape(-1, 3)
        abi.check(self.lib, self.lib.nsb_cage_deform(self.ctx, op_index, cv.ctypes.data, cv.shape[0], stream), "nsb_cage_deform")

    def cage_download(self, op_index: int, n_vertices: int, n_tets: int, rotations: bool = True) -> dict:
        n = C.c_uint64()
        abi.check(self.lib, self.lib.nsb_cage_download(self.ctx, op_index, None, None, None, None, 0, C.byref(n), None), "nsb_cage_download")
        out = dict(vertices=np.zeros((n_vertices, 3), np.float32), rotations=np.zeros((n_tets, 9), np.float32) if rotations else None,
                   lut_offsets=np.zeros(abi.NSB_GRID_CELLS + 1, np.uint32), lut_idx=np.zeros(max(int(n.value), 1), np.uint32), boxes=np.zeros(12, np.float32))
        abi.check(self.lib, self.lib.nsb_cage_download(self.ctx, op_index, out["vertices"].ctypes.data, out["rotations"].ctypes.data if rotations else None,
                                                       out["lut_offsets"].ctypes.data, out["lut_idx"].ctypes.data, out["lut_idx"].size, C.byref(n),
                                                       out["boxes"].ctypes.data), "nsb_cage_download")
        out["lut_idx"] = out["lut_idx"][: int(n.value)]
        return out

    # ---- membrane boundary values (GrowingSelection::compute_poisson_boundary / interpolate_poisson_boundary) ----
    def poisson_boundary(self, points: np.ndarray, is_inside: bool, seed: int = 0, sampling_width: int = 10, hemisphere_width: int = 10,
                         train_aabb=((-1.5, -1.5, -1.5), (2.5, 2.5, 2.5)), rgb_activation: int = abi.NSB_ACT_LOGISTIC,
                         density_activation: int = abi.NSB_ACT_EXPONENTIAL):
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        p = abi.NsbBoundarySampling()
        p.sampling_width, p.hemisphere_width, p.seed = sampling_width, hemisphere_width, seed
        p.train_aabb_min[:] = train_aabb[0]
        p.train_aabb_max[:] = train_aabb[1]
        p.rgb_activation, p.density_activation, p.is_inside = rgb_activation, density_activation, int(is_inside)
        dens, shs = np.zeros(pts.shape[0], np.float32), np.zeros((pts.shape[0], 27), np.float32)
        abi.check(self.lib, self.lib.nsb_poisson_boundary(self.ctx, pts.ctypes.data, pts.shape[0], C.byref(p), dens.ctypes.data, shs.ctypes.data), "nsb_poisson_boundary")
        return dens, shs, p

    def cage_set_membrane(self, op_index: int, gamma: np.ndarray, inside_density, outside_density, inside_shs, outside_shs, amplitude: float = 1.0,
                          apply: bool = True):
        g = np.ascontiguousarray(gamma, np.float32)
        arrs = [np.ascontiguousarray(a, np.float32) for a in (inside_density, outside_density, inside_shs, outside_shs)]
        nv = g.shape[0]
        b_shs, b_od, b_rd = np.zeros((nv, 27), np.float32), np.zeros(nv, np.float32), np.zeros(nv, np.float32)
        abi.check(self.lib, self.lib.nsb_cage_set_membrane(self.ctx, op_index, g.ctypes.data, g.shape[1], *[a.ctypes.data for a in arrs], float(amplitude), int(apply),
                                                           b_shs.ctypes.data, b_od.ctypes.data, b_rd.ctypes.data), "nsb_cage_set_membrane")
        return b_shs, b_od, b_rd

    def set_edit_operators(self, ops):
        """ops: list of (NsbEditOp, keepalive) in m_edit_operators order."""
        self._ops = list(ops or [])
        if self._ops:
            arr = (abi.NsbEditOp * len(self._ops))(*[o[0] for o in self._ops])
            abi.check(self.lib, self.lib.nsb_set_edit_ops(self.ctx, arr, len(self._ops)), "nsb_set_edit_ops")
        else:
            abi.check(self.lib, self.lib.nsb_set_edit_ops(self.ctx, None, 0), "nsb_set_edit_ops")

    def add_edit_operator(self, op):
        self.set_edit_operators(self._ops + [op])

    def reset_edit_operators(self):
        self.set_edit_operators([])

    # ---- the hot path ------------------------------------------------------------------------------------
    def render(self, frame: abi.NsbFrame, fb=None, depth=None, stream=None):
        """fb: float32 CUDA tensor [H, W, 4] (over-blended onto, like render_buffer.frame_buffer()); depth [H, W]."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        if fb is None:
            fb = torch.zeros((frame.height, frame.width, 4), dtype=torch.float32, device=dev)
        if depth is None:
            depth = torch.zeros((frame.height, frame.width), dtype=torch.float32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        abi.check(self.lib, self.lib.nsb_render(self.ctx, C.byref(frame), fb.data_ptr(), depth.data_ptr(), s), "nsb_render")
        return fb, depth

    def render_to_cpu(self, frame: abi.NsbFrame, fb_host=None, depth_host=None):
        """Host-buffer entry point (pinned torch tensors or numpy arrays)."""
        if fb_host is None:
            fb_host = np.zeros((frame.height, frame.width, 4), np.float32)
        if depth_host is None:
            depth_host = np.zeros((frame.height, frame.width), np.float32)
        fp = fb_host.data_ptr() if hasattr(fb_host, "data_ptr") else fb_host.ctypes.data
        dp = depth_host.data_ptr() if hasattr(depth_host, "data_ptr") else depth_host.ctypes.data
        abi.check(self.lib, self.lib.nsb_render_host(self.ctx, C.byref(frame), fp, dp), "nsb_render_host")
        return fb_host, depth_host

    def render_to_cpu_async(self, frame: abi.NsbFrame, fb_host, depth_host=None) -> int:
        """nsb_render_host_async: returns a ticket; fb_host / depth_host (pinned) are valid after wait_host_frame(ticket). The copy of this frame
        overlaps the render of the next one."""
        fp = fb_host.data_ptr() if hasattr(fb_host, "data_ptr") else fb_host.ctypes.data
        dp = None if depth_host is None else (depth_host.data_ptr() if hasattr(depth_host, "data_ptr") else depth_host.ctypes.data)
        ticket = C.c_uint64()
        abi.check(self.lib, self.lib.nsb_render_host_async(self.ctx, C.byref(frame), fp, dp, C.byref(ticket)), "nsb_render_host_async")
        return ticket.value

    def wait_host_frame(self, ticket: int) -> None:
        abi.check(self.lib, self.lib.nsb_host_frame_wait(self.ctx, ticket), "nsb_host_frame_wait")

    def stats(self) -> abi.NsbRenderStats:
        st = abi.NsbRenderStats()
        abi.check(self.lib, self.lib.nsb_get_stats(self.ctx, C.byref(st)), "nsb_get_stats")
        return st

    def debug_counters(self) -> dict:
        out = (C.c_uint64 * 7)()
        abi.check(self.lib, self.lib.nsb_debug_counters(self.ctx, out, 7), "nsb_debug_counters")
        keys = ("rounds", "cyc_acquire", "cyc_encode", "cyc_mlp", "cyc_composite", "cyc_total", "ctas")
        return dict(zip(keys, [int(v) for v in out]))

    # ---- multi-GPU framebuffer shards ------------------------------------------------------------------
    def tiles_for_rank(self, width, height, rank, world) -> int:
        n = C.c_uint32()
        abi.check(self.lib, self.lib.nsb_tiles_for_rank(width, height, rank, world, C.byref(n)), "nsb_tiles_for_rank")
        return n.value

    def pack_tiles(self, fb, depth, rank, world, out_rgba, out_depth=None):
        torch = _torch()
        H, W = fb.shape[0], fb.shape[1]
        s = torch.cuda.current_stream().cuda_stream
        abi.check(self.lib, self.lib.nsb_pack_tiles(self.ctx, fb.data_ptr(), 0 if depth is None else depth.data_ptr(), W, H, rank, world,
                                                    out_rgba.data_ptr(), 0 if out_depth is None else out_depth.data_ptr(), s), "nsb_pack_tiles")

    def unpack_tiles(self, packed_rgba, packed_depth, rank, world, fb, depth=None):
        torch = _torch()
        H, W = fb.shape[0], fb.shape[1]
        s = torch.cuda.current_stream().cuda_stream
        abi.check(self.lib, self.lib.nsb_unpack_tiles(self.ctx, packed_rgba.data_ptr(), 0 if packed_depth is None else packed_depth.data_ptr(), W, H, rank, world,
                                                      fb.data_ptr(), 0 if depth is None else depth.data_ptr(), s), "nsb_unpack_tiles")

    def unpack_gathered(self, gathered_rgba, gathered_depth, rank_stride_floats, skip_rank, world, fb, depth=None):
        """ONE launch: every rank's packed tiles of an all-gather result (rank k's float4 tiles at gathered_rgba + k*stride, its depth floats at
        gathered_depth + k*stride) except skip_rank's -> fb (+ depth)."""
        torch = _torch()
        H, W = fb.shape[0], fb.shape[1]
        s = torch.cuda.current_stream().cuda_stream
        abi.check(self.lib, self.lib.nsb_unpack_gathered(self.ctx, gathered_rgba.data_ptr(), 0 if gathered_depth is None else gathered_depth.data_ptr(), W, H, skip_rank, world,
                                                         int(rank_stride_floats), fb.data_ptr(), 0 if depth is None else depth.data_ptr(), s), "nsb_unpack_gathered")

    # ---- frame post-process (CudaRenderBuffer::accumulate / ::tonemap) ---------------------------------------
    def accumulate(self, frame_buffer, accumulate_buffer, spp: int, color_space: int = abi.NSB_COLOR_LINEAR):
        torch = _torch()
        H, W = frame_buffer.shape[0], frame_buffer.shape[1]
        abi.check(self.lib, self.lib.nsb_accumulate(self.ctx, frame_buffer.data_ptr(), accumulate_buffer.data_ptr(), W, H, spp, color_space,
                                                    torch.cuda.current_stream().cuda_stream), "nsb_accumulate")

    def tonemap(self, accumulate_buffer, params: "abi.NsbTonemap", out=None):
        torch = _torch()
        H, W = accumulate_buffer.shape[0], accumulate_buffer.shape[1]
        if out is None:
            out = torch.empty_like(accumulate_buffer)
        abi.check(self.lib, self.lib.nsb_tonemap(self.ctx, accumulate_buffer.data_ptr(), out.data_ptr(), W, H, C.byref(params),
                                                 torch.cuda.current_stream().cuda_stream), "nsb_tonemap")
        return out

    # ---- operator-level entry points ---------------------------------------------------------------------
    def _coords(self, coords):
        torch = _torch()
        c = torch.as_tensor(np.ascontiguousarray(coords, np.float32)) if not hasattr(coords, "data_ptr") else coords
        return c.to(torch.device("cuda", self.device)).contiguous()

    def _run_net(self, fn, name, coords, rows):
        torch = _torch()
        c = self._coords(coords)
        n = c.shape[0]
        n_pad = ((n + 127) // 128) * 128
        if n == 0:
            abi.check(self.lib, fn(self.ctx, None, 0, None, 0, None), name)  # empty batch is a no-op
            return np.zeros((rows, 0), np.uint16)
        out = torch.zeros((rows, n_pad), dtype=torch.int16, device=c.device)
        abi.check(self.lib, fn(self.ctx, c.data_ptr(), n, out.data_ptr(), n_pad, torch.cuda.current_stream().cuda_stream), name)
        torch.cuda.synchronize()
        return out.cpu().numpy().view(np.uint16)[:, :n]

    def inference(self, coords) -> np.ndarray:
        return self._run_net(self.lib.nsb_inference, "nsb_inference", coords, 16)

    def density(self, coords) -> np.ndarray:
        return self._run_net(self.lib.nsb_density, "nsb_density", coords, 16)

    def encode(self, coords) -> np.ndarray:
        return self._run_net(self.lib.nsb_encode, "nsb_encode", coords, 32)

    def map_rays(self, coords):
        torch = _torch()
        c = self._coords(coords).clone()
        n = c.shape[0]
        mask = torch.zeros(n, dtype=torch.uint8, device=c.device)
        abi.check(self.lib, self.lib.nsb_map_rays(self.ctx, c.data_ptr(), mask.data_ptr(), n, torch.cuda.current_stream().cuda_stream), "nsb_map_rays")
        torch.cuda.synchronize()
        return c.cpu().numpy(), mask.cpu().numpy()

    def poisson_residuals(self, coords):
        torch = _torch()
        c = self._coords(coords)
        n = c.shape[0]
        sh = torch.zeros((n, 27), dtype=torch.float32, device=c.device)
        od = torch.zeros(n, dtype=torch.float32, device=c.device)
        rd = torch.zeros(n, dtype=torch.float32, device=c.device)
        abi.check(self.lib, self.lib.nsb_poisson_residuals(self.ctx, c.data_ptr(), n, sh.data_ptr(), od.data_ptr(), rd.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream), "nsb_poisson_residuals")
        torch.cuda.synchronize()
        return sh.cpu().numpy(), od.cpu().numpy(), rd.cpu().numpy()

    def march_trace(self, frame: abi.NsbFrame, pixels: np.ndarray, max_samples: int):
        torch = _torch()
        dev = torch.device("cuda", self.device)
        p = torch.as_tensor(np.ascontiguousarray(pixels, np.uint32).view(np.int32)).to(dev)
        n = p.numel()
        rec = torch.zeros((n, max_samples, 5), dtype=torch.float32, device=dev)
        idx = torch.zeros((n, max_samples, 2), dtype=torch.int32, device=dev)
        cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        abi.check(self.lib, self.lib.nsb_march_trace(self.ctx, C.byref(frame), p.data_ptr(), n, max_samples, rec.data_ptr(), idx.data_ptr(), cnt.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "nsb_march_trace")
        torch.cuda.synchronize()
        return rec.cpu().numpy(), idx.cpu().numpy().view(np.uint32), cnt.cpu().numpy().view(np.uint32)
