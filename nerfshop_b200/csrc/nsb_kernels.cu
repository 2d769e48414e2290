// nsb_kernels.cu — kernels of the B200-native NeRFshop render path and the C ABI on top of them.
//
// The product kernel is k_render_fused: ONE persistent launch per frame that replaces the reference's
// per-round pipeline (compact -> generate inputs -> [residuals] -> inference -> [map_rays] -> inference ->
// composite, >= 12 launches + 3 host syncs per round, testbed_nerf.cu:2812-2990) and its 4.6 GB of
// scratch: ray state lives in registers and shared memory, samples go march -> deform -> hash encode -> tcgen05 MLP ->
// composite without touching HBM, and finished rays are shaded straight into the framebuffer.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/nerfshop_b200.h"
#include "nsb_device.cuh"
#include "nsb_tc.cuh"

using namespace nsb;

// =====================================================================================================
// fused persistent renderer
// =====================================================================================================
struct RayRec { uint32_t pix; float t; };  // a ray that entered the AABB and found occupied space: pixel + first sample t

enum { ST_RAYS = 0, ST_ALIVE = 1, ST_HIT = 2, ST_SAMPLES = 3, ST_OLD = 4, ST_ROUNDS = 5, ST_CYC_ACQUIRE = 6, ST_CYC_ENCODE = 7, ST_CYC_MLP = 8,
       ST_CYC_COMPOSITE = 9, ST_CYC_TOTAL = 10, ST_CTAS = 11, ST_N = 12 };
constexpr int DDA_BUDGET = 8;   // empty voxels a lane may skip per MLP round before it yields (latency bound of a round)

// tile slot -> pixel: the four warps of a 16x8 tile each cover an 8x4 block so that the 32 lanes of a warp stay
// spatially compact (coherent hash-grid and occupancy lookups)
__device__ __forceinline__ bool tile_pixel(const DevFrame& f, uint32_t tile, uint32_t slot, uint32_t& px, uint32_t& py) {
	uint32_t tx = tile % (uint32_t)f.tiles_x, ty = tile / (uint32_t)f.tiles_x;
	uint32_t b = slot >> 5, l = slot & 31u;
	px = tx * TILE_W + (b & 1u) * 8u + (l & 7u);
	py = ty * TILE_H + (b >> 1) * 4u + (l >> 3);
	return px < (uint32_t)f.W && py < (uint32_t)f.H;
}

// Stage 1 (massively parallel, one thread per pixel of this rank's tiles):
// init_rays_with_payload_kernel_nerf (testbed_nerf.cu:2512) + advance_pos_nerf (:557) + the alive-ray compaction
// of compact_kernel_nerf (:2485), with warp-aggregated atomics instead of one global atomic per ray.
// The long empty-space walk to a ray's first occupied sample is a serial chain of dependent occupancy loads;
// here it is hidden by running every ray of the frame concurrently instead of stalling an MLP round.
// GENERAL: frames with lens distortion / a distortion map / depth of field (f.general_camera), an environment map or the Distortion render mode.
// Their rays cost a Newton iteration or a sincosf to generate, so the queued ones are handed to the persistent renderer in `ray_od` (origin and
// direction, 6 floats per queue slot) instead of being re-derived from the pixel index there; the common pinhole path stays as lean as it was.
template <bool GENERAL>
__global__ void __launch_bounds__(256) k_prepare_rays(const DevFrame f, const uint8_t* __restrict__ bitfield, float4* __restrict__ fb, float* __restrict__ depth_out,
                                                      RayRec* __restrict__ list, float* __restrict__ ray_od, uint32_t* __restrict__ n_queued, uint32_t n_local_pixels,
                                                      unsigned long long* __restrict__ stats) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	bool queued = false, is_ray = false, entered = false;
	uint32_t pix = 0;
	float t = 0.0f;
	Ray r;
	if (i < n_local_pixels) {
		uint32_t tile = (uint32_t)f.tile_rank + (i / (uint32_t)TILE_PIXELS) * (uint32_t)f.tile_world;
		uint32_t px, py;
		if (tile_pixel(f, tile, i % (uint32_t)TILE_PIXELS, px, py)) {
			is_ray = true;
			pix = px + (uint32_t)f.W * py;
			depth_out[pix] = 1e10f;  // :2581
			bool inside;
			if constexpr (GENERAL) {
				inside = f.general_camera ? make_ray_general(f, px, py, r) : make_ray(f, px, py, r);
				if (f.envmap) fb[pix] = read_envmap(f.envmap, f.env_w, f.env_h, r.d);  // every pixel, before the AABB test (:2581-2583)
				if (inside && f.mode == NSB_RENDER_DISTORTION) {  // :2596-2607: the map itself, written by ray initialisation; the ray is not traced
					float dx = 0.5f, dy = 0.5f;
					if (f.distortion) {
						read_image2(f.distortion, f.dist_w, f.dist_h, ((float)px + 0.5f) / (float)f.W, ((float)py + 0.5f) / (float)f.H, dx, dy);
						dx = dx * 50.0f + 0.5f; dy = dy * 50.0f + 0.5f;
					}
					fb[pix] = make_float4(dx, dy, 0.5f, 1.0f);
					depth_out[pix] = 1.0f;
					inside = false;
				}
			} else {
				inside = make_ray(f, px, py, r);
			}
			if (inside) {
				entered = true;
				V3 idir = v3(div_(1.0f, r.d.x), div_(1.0f, r.d.y), div_(1.0f, r.d.z));
				t = fma_(ld_random_val(f.spp, pix * 786433u), calc_dt(r.t, f.cone), r.t);  // advance_pos_nerf :585
				float dt; V3 pos; uint32_t mip, cell;
				queued = next_occupied(f, bitfield, r.o, r.d, idir, t, dt, pos, mip, cell);
			}
		}
	}
	const unsigned lane = threadIdx.x & 31u;
	unsigned m = __ballot_sync(0xffffffffu, queued);
	unsigned mr = __ballot_sync(0xffffffffu, is_ray), me = __ballot_sync(0xffffffffu, entered);
	uint32_t base = 0;
	if (lane == 0) {
		if (m) base = atomicAdd(n_queued, (uint32_t)__popc(m));
		if (mr) atomicAdd(stats + ST_RAYS, (unsigned long long)__popc(mr));
		if (me) atomicAdd(stats + ST_ALIVE, (unsigned long long)__popc(me));
	}
	base = __shfl_sync(0xffffffffu, base, 0);
	if (queued) {
		RayRec rr;
		rr.pix = pix;
		rr.t = t;
		const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
		list[slot] = rr;
		if constexpr (GENERAL) {
			if (ray_od) {
				float* od = ray_od + 6 * (size_t)slot;
				od[0] = r.o.x; od[1] = r.o.y; od[2] = r.o.z; od[3] = r.d.x; od[4] = r.d.y; od[5] = r.d.z;
			}
		}
	}
}

// grid features of one sample -> this thread's row of the A operand (4 chunks x 8 fp16)
// Code size matters here: the fused kernel's hot loop has to live in the SM's instruction cache (profiles/README.md:
// the first, fully unrolled version was 504 KB of SASS and stalled on instruction fetch). One loop body = 2 levels
// (16 gathers in flight per thread), 8 iterations.
#ifndef NSB_ENC_NL
#define NSB_ENC_NL 2  // levels per loop iteration (8*NL gathers in flight per thread)
#endif
__device__ __forceinline__ void encode_to_a32(uint8_t* a32, const DevModel& m, bool valid, V3 pw, uint32_t row) {
	constexpr int NL = NSB_ENC_NL;
#pragma unroll 1
	for (int g = 0; g < MAX_LEVELS / NL; ++g) {
		__half2 h[NL];
#pragma unroll
		for (int j = 0; j < NL; ++j) h[j] = __floats2half2_rn(0.0f, 0.0f);
		if (valid) {
			uint32_t mode = m.pair_mode[(NL * g) / 2];  // uniform
			if (NL == 4 && m.pair_mode[(NL * g) / 2 + 1] != mode) mode = 2;
			if (mode == 1) encode_levels<NL, 1>(m.levels + NL * g, m.grid, pw.x, pw.y, pw.z, h);
			else if (mode == 0) encode_levels<NL, 0>(m.levels + NL * g, m.grid, pw.x, pw.y, pw.z, h);
			else encode_levels<NL, 2>(m.levels + NL * g, m.grid, pw.x, pw.y, pw.z, h);
		}
		// levels NL*g .. NL*g+NL-1 = fp16 features 2*NL*g ..: NL*4 bytes of this row's k-chunk (NL*g)/4
		uint8_t* dst = a32 + ((NL * g) >> 2) * (tc::ROWS * 16) + row * 16 + ((NL * g) & 3) * 4;
		if (NL == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(tc::pack_h2(h[0]), tc::pack_h2(h[1]));
		else *reinterpret_cast<uint4*>(dst) = make_uint4(tc::pack_h2(h[0]), tc::pack_h2(h[1]), tc::pack_h2(h[NL > 2 ? 2 : 0]), tc::pack_h2(h[NL > 3 ? 3 : 0]));
	}
}

__device__ __forceinline__ float h_lo(uint32_t packed) { return __half2float(__ushort_as_half((unsigned short)(packed & 0xffffu))); }
__device__ __forceinline__ float h_hi(uint32_t packed) { return __half2float(__ushort_as_half((unsigned short)(packed >> 16))); }

// position of the k-th (0-based) set bit of m; popc(m) > k
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t m, uint32_t k) {
	uint32_t p = 0;
#pragma unroll
	for (uint32_t s = 16; s >= 1; s >>= 1) {
		const uint32_t c = (uint32_t)__popc((m >> p) & ((1u << s) - 1u));
		if (k >= c) { k -= c; p += s; }
	}
	return p;
}

// Stage 2 (persistent; a tile = 128 threads = 128 ray slots = 128 MMA rows, NSB_TILES tiles per CTA): every round each live ray contributes
// ONE sample; samples go deform -> hash encode -> tcgen05 MLPs -> composite inside the SM; a finished ray is shaded
// into the framebuffer and its slot refilled from the queue (warp-convergent fetch of consecutive entries, so rays that
// were refilled together — neighbouring pixels that die on the same surface — stay together in a warp).
// Per-slot ray state parked in shared memory (SoA: lane-contiguous, conflict-free). Only the acquire and composite phases
// touch it, so it does not occupy registers while the thread gathers hash-grid features and walks the MLP layers.
enum { R_OX = 0, R_OY, R_OZ, R_DX, R_DY, R_DZ, R_T, R_CR, R_CG, R_CB, R_CA, R_DEPTH, R_MAXW, R_FIELDS };
// One CTA per SM holds NSB_TILES independent tiles (a tile = 128 threads = 128 ray slots = one UMMA M). The tiles share the 20 KB
// weight image — four separate CTAs would hold four copies — which matters because shared memory is carved out of the same
// 256 KB as the L1 that serves the hash-grid gathers: 4 x 43.6 KB CTAs leave ~60 KB of L1, one 4-tile CTA (112 KB) leaves ~124 KB
// (profiles/README.md: the tile loop runs 44 % faster with 92 KB of L1 than with 28 KB).
// How many: as many as the register file allows without spilling. Frames without operators need 80 registers per thread -> 6 tiles (24 warps per SM,
// 158 KB of shared memory, ~97 KB of L1): 10.96 ms per 1080p frame against 12.63 ms with 4 tiles and 11.51 ms with 5 (round 2, profiles/README.md
// item 23; round 1's fatter loop spilled beyond 4). Frames with operators carry the deform / membrane state: 5 tiles (96 registers, 40 B spilled).
#ifndef NSB_TILES
#define NSB_TILES 6
#endif
#ifndef NSB_TILES_OPS
#define NSB_TILES_OPS 5
#endif
struct TileBlock {
	union {
		uint8_t a64[tc::A64_BYTES];
		uint8_t a32[tc::A32_BYTES];
	};
	float ray[R_FIELDS][128];
	uint64_t mma_bar;
	uint64_t pad[15];
};
template <int TILES>
struct __align__(128) RenderSmemT {
	uint8_t w[tc::W_BYTES];
	TileBlock tile[TILES];
	uint64_t w_bar;
	uint32_t tmem_base;
	uint32_t pad;
};
__host__ __device__ constexpr int render_tiles(bool ops) { return ops ? NSB_TILES_OPS : NSB_TILES; }
__host__ __device__ constexpr uint32_t render_tmem_cols(int tiles) { return tiles <= 1 ? 64u : tiles <= 2 ? 128u : tiles <= 4 ? 256u : 512u; }
static_assert(NSB_TILES >= 1 && NSB_TILES <= 8 && NSB_TILES_OPS >= 1 && NSB_TILES_OPS <= 8, "64 TMEM columns per tile, 512 per SM");

// OPS = false: the instantiation for frames without edit operators (f.apply_ops == 0 or no operator uploaded) carries none of the deform /
// membrane code in its hot loop (the loop has to fit the instruction cache: profiles/README.md item 6).
// ACC16: the MLP accumulator policy (nsb_set_mlp_accumulator): fp16 TMEM accumulators, like the reference's wmma __half fragments.
// experiment switches (tools/ab_kernels.py variants): compile a feature of the fused kernel out to measure what its presence costs the default path
#ifndef NSB_X_ACCEL
#define NSB_X_ACCEL 1
#endif
#ifndef NSB_X_GLOW
#define NSB_X_GLOW 1
#endif
#ifndef NSB_X_RAYOD
#define NSB_X_RAYOD 1
#endif
#ifndef NSB_X_HELP
#define NSB_X_HELP 0  // helper lanes: measured slower when used AND 5 % slower by their mere presence in the loop (register allocation / code layout):
                      // compiled out of the product; `build.build_variant("helpers", ["NSB_X_HELP=1"])` is the experiment build NSB_HELPERS acts on
#endif
template <bool OPS, bool ACC16>
__global__ void __launch_bounds__(128 * render_tiles(OPS), 1) k_render_fused(const DevFrame f, const DevModel m, const uint8_t* __restrict__ bitfield,
                                                      const DevOp* __restrict__ ops, const int n_ops, const int any_poisson,
                                                      float4* __restrict__ fb, float* __restrict__ depth_out, const RayRec* __restrict__ list,
                                                      const float* __restrict__ ray_od, const uint32_t* __restrict__ n_queued_ptr, uint32_t* fetch_counter,
                                                      unsigned long long* __restrict__ stats, const int refill_thr, const int dda_flags) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	constexpr int TILES = render_tiles(OPS);
	constexpr uint32_t RENDER_TMEM_COLS = render_tmem_cols(TILES);
	using RenderSmem = RenderSmemT<TILES>;
	RenderSmem& RS = *reinterpret_cast<RenderSmem*>(smem_raw);
	const uint32_t tid = threadIdx.x & 127u;      // slot / row within the tile
	const uint32_t tile = threadIdx.x >> 7;
	TileBlock& TB = RS.tile[tile];
	const uint32_t lane = tid & 31u;
	const uint32_t n_queued = *n_queued_ptr;
	if (blockIdx.x * 128u * TILES >= n_queued && blockIdx.x > 0) return;  // nothing this CTA could ever fetch

	// CTA setup: barriers, TMEM (64 columns per tile), one bulk-TMA copy of the weight image
	if (threadIdx.x == 0) {
		tc::mbar_init(&RS.w_bar, 1);
		for (int t = 0; t < TILES; ++t) tc::mbar_init(&RS.tile[t].mma_bar, 1);
		tc::fence_mbar_init();
	}
	if (threadIdx.x < 32) tc::tmem_alloc(&RS.tmem_base, RENDER_TMEM_COLS);
	tc::tc_fence_before();
	__syncthreads();
	tc::tc_fence_after();
	if (threadIdx.x < 32) {
		if (threadIdx.x == 0) {
			tc::mbar_expect_tx(&RS.w_bar, tc::W_BYTES);
			tc::bulk_g2s(RS.w, m.w_image, tc::W_BYTES, &RS.w_bar);
		}
		__syncwarp();
	}
	tc::mbar_wait(&RS.w_bar, 0);
	const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&RS.tmem_base);
	tc::TileCtx C;
	C.a32 = TB.a32; C.a64 = TB.a64; C.w_addr = tc::smem_u32(RS.w); C.mma_bar = &TB.mma_bar; C.tmem = tmem_base + tile * tc::TMEM_COLS; C.row = tid;
	C.bar_id = 1 + tile;

	const int dda_budget = dda_flags & 0xffff;
	// Helper lanes (frames without operators). A round costs the same ~20 K cycles whether 32 or 3 lanes of a warp carry a sample, and a warp only
	// refills once it is down to refill_thr rays, so the lanes WITHOUT a ray are lent to the rays that found a sample this round: each such ray gets up
	// to HELP_DEPTH of them, and helper j evaluates the ray's j-th NEXT sample (t advanced j times by calc_dt, its own occupancy test) through the
	// encode and the MLPs. The owner lane then composites its own sample and, in march order, the helpers' — up to the first one that was not an
	// occupied in-bounds sample, where the ordinary march resumes next round. Samples past a ray's termination are evaluated and discarded, like the
	// reference's own n_steps > 1 batching; frames are bit-identical with or without helpers. help_mode 1: only once the queue is exhausted, 2: always.
	// MEASURED (profiles/README.md, round 2): 13.5 % fewer rounds but each round 27 % dearer — the loop's cost follows the samples evaluated (gather
	// wavefronts + issue), not the rounds — and the code's presence alone costs 5 %: it is compiled only into the experiment build (NSB_X_HELP=1).
	const bool fx_show_accel = NSB_X_ACCEL && f.show_accel;
	const int fx_glow = NSB_X_GLOW ? f.glow_mode : 0;
	const float* const fx_ray_od = NSB_X_RAYOD ? ray_od : nullptr;
	const int help_mode = (!NSB_X_HELP || OPS || (fx_glow & 4)) ? 0 : (dda_flags >> 16) & 3;  // host: NSB_HELPERS (the mailbox does not carry glow's weight mask)
	constexpr uint32_t HELP_DEPTH = 7;
	// mailbox helper -> owner: the tile's A-operand buffer is idle between the last MMA of a round and the next encode, and every warp only ever
	// writes its own 32 rows of it: field q of lane l lives in k-chunk q, at this warp's 512 bytes
	float* const mbox = reinterpret_cast<float*>(TB.a64 + (tid >> 5) * 512u) + lane;
	constexpr uint32_t MB = tc::ROWS * 16 / 4;  // floats between mailbox fields
	enum { M_ALPHA = 0, M_E0, M_E1, M_E2, M_Z, M_TNEXT };  // M_TNEXT < 0: not a sample
	constexpr bool ops_on = OPS;  // the host picks the instantiation: f.apply_ops && n_ops > 0

	// what stays in registers across a round
	bool alive = false, exhausted = false;  // exhausted is warp-uniform: the queue had nothing left when this warp last asked
	uint32_t ch_next = 0, ch_end = 0;       // this warp's chunk of the ray queue (warp-uniform)
	uint32_t pix = 0, n_steps = 0;
	uint32_t c_hit = 0, c_samples = 0, c_old = 0;
	uint32_t phase = 0;
	// phase profile (thread 0 of each CTA; a handful of clock reads per round): diagnostic builds only (-DNSB_PROFILE), the
	// product build keeps the registers and issue slots (nsb_debug_counters then reports rounds and CTAs only)
#ifdef NSB_PROFILE
#define NSB_CLOCK() clock64()
	long long cyc_acq = 0, cyc_enc = 0, cyc_mlp = 0, cyc_comp = 0;
	const long long cyc_start = clock64();
#else
#define NSB_CLOCK() 0ll
#endif
	uint32_t n_rounds = 0;

	// shade_kernel_nerf (:2464-2482) for the ray this thread just finished (compact_kernel_nerf :2503 filter)
	auto finish = [&](bool left_aabb, float cr, float cg, float cb, float ca, float ray_depth) {
		alive = false;
		if (!(ca > 0.001f)) return;
		++c_hit;
		float r = cr, g = cg, b = cb, a = ca;
		if (f.mode == NSB_RENDER_COST) {
			float col = (float)(n_steps + (left_aabb ? 1u : 0u)) / 128.0f;
			r = g = b = col;
			a = 1.0f;
		}
		if (!f.linear_colors && f.mode == NSB_RENDER_SHADE) { r = srgb_to_linear(r); g = srgb_to_linear(g); b = srgb_to_linear(b); }
		float4 prev = fb[pix];
		float om = 1.0f - a;
		fb[pix] = make_float4(__fmaf_rn(prev.x, om, r), __fmaf_rn(prev.y, om, g), __fmaf_rn(prev.z, om, b), __fmaf_rn(prev.w, om, a));
		if (a > 0.2f) depth_out[pix] = ray_depth;
	};

	for (;;) {
		if (!tc::tile_any(C, alive || !exhausted)) break;
#ifdef NSB_PROFILE
		const long long c0 = clock64();
#endif

		// ---- acquire one occupied sample for this thread (refill the ray slot when it is free) ----
		bool has_sample = false;
		float dt = 0.0f;
		V3 pos = v3(0, 0, 0), dw = v3(0.5f, 0.5f, 0.5f);
		int h_owner = -1;                  // >= 0: this lane helps the ray of that lane this round
		float h_tnext = -1.0f;             // helper: t after its sample
		unsigned h_free = 0u;              // warp-uniform: the helper lanes of this round (0: none)
		uint32_t h_per_ray = 0, h_first = 0;  // owner: its helpers are the h_per_ray free lanes from free-rank h_first
		{
			int budget = dda_budget;
			V3 ro = v3(TB.ray[R_OX][tid], TB.ray[R_OY][tid], TB.ray[R_OZ][tid]);
			V3 rd = v3(TB.ray[R_DX][tid], TB.ray[R_DY][tid], TB.ray[R_DZ][tid]);
			float t = TB.ray[R_T][tid];
			// Refill (warp-convergent). Free lanes wait until the warp is down to <= refill_thr live rays; rays come from a queue CHUNK
			// this warp owns (consecutive entries = neighbouring pixels), so the lanes of a warp gather from the same few hash-grid
			// cells on the coarse and middle levels instead of 32 unrelated cache lines per instruction.
			{
				const unsigned alive_m = __ballot_sync(0xffffffffu, alive);
				const unsigned want = exhausted ? 0u : ~alive_m;
				if (want && __popc(alive_m) <= (refill_thr & 0xff)) {
					const uint32_t need = (uint32_t)__popc(want), rank = (uint32_t)__popc(want & ((1u << lane) - 1u));
					const uint32_t chunk = (uint32_t)refill_thr >> 8;
					uint32_t given = 0, qi = 0xffffffffu;
					while (given < need) {  // warp-uniform
						if (ch_next >= ch_end) {
							const uint32_t claim = chunk ? chunk : need - given;
							uint32_t base = 0;
							if (lane == 0) base = atomicAdd(fetch_counter, claim);
							base = __shfl_sync(0xffffffffu, base, 0);
							if (base >= n_queued) { exhausted = true; break; }
							ch_next = base;
							ch_end = min(base + claim, n_queued);
						}
						const uint32_t take = min(need - given, ch_end - ch_next);
						if (!alive && rank >= given && rank < given + take) qi = ch_next + (rank - given);
						ch_next += take;
						given += take;
					}
					if (qi != 0xffffffffu) {
						const RayRec rr = list[qi];
						pix = rr.pix;
						if (fx_ray_od) {  // a general-camera frame: k_prepare_rays<true> left the ray next to its queue entry
							const float* od = fx_ray_od + 6 * (size_t)qi;
							ro = v3(od[0], od[1], od[2]); rd = v3(od[3], od[4], od[5]);
						} else {
							Ray r;
							make_ray(f, pix % (uint32_t)f.W, pix / (uint32_t)f.W, r);
							ro = r.o; rd = r.d;
						}
						t = rr.t;
						TB.ray[R_OX][tid] = ro.x; TB.ray[R_OY][tid] = ro.y; TB.ray[R_OZ][tid] = ro.z;
						TB.ray[R_DX][tid] = rd.x; TB.ray[R_DY][tid] = rd.y; TB.ray[R_DZ][tid] = rd.z;
						TB.ray[R_CR][tid] = 0.0f; TB.ray[R_CG][tid] = 0.0f; TB.ray[R_CB][tid] = 0.0f; TB.ray[R_CA][tid] = 0.0f;
						TB.ray[R_DEPTH][tid] = 0.0f; TB.ray[R_MAXW][tid] = 0.0f;
						n_steps = 0;
						alive = true;
					}
				}
			}
			for (;;) {
				if (!alive) break;
				if (n_steps >= MARCH_ITER - 1) { alive = false; continue; }  // still marching after MARCH_ITER steps: dropped (:2812)
				const V3 idir = v3(div_(1.0f, rd.x), div_(1.0f, rd.y), div_(1.0f, rd.z));
				const MarchResult mr = next_occupied_budget(f, bitfield, ro, rd, idir, t, dt, pos, budget);
				if (mr == MARCH_FOUND) { has_sample = true; break; }
				if (mr == MARCH_EXIT) {
					finish(true, TB.ray[R_CR][tid], TB.ray[R_CG][tid], TB.ray[R_CB][tid], TB.ray[R_CA][tid], TB.ray[R_DEPTH][tid]);
					continue;
				}
				break;  // MARCH_PENDING: resume next round
			}
			if (has_sample) {
				dw = warp_direction(rd);
				t = add(t, dt);
			}
			if (alive) TB.ray[R_T][tid] = t;
			if constexpr (!OPS) {
				if (help_mode == 2 || (help_mode == 1 && exhausted)) {
					const unsigned smp_m = __ballot_sync(0xffffffffu, has_sample);  // rays that found a sample this round
					const unsigned free_m = ~__ballot_sync(0xffffffffu, alive);     // lanes without a ray
					const uint32_t n_smp = (uint32_t)__popc(smp_m), n_free = (uint32_t)__popc(free_m);
					if (n_smp != 0u && n_free != 0u) {
						h_free = free_m;
						h_per_ray = n_free >= n_smp ? min(HELP_DEPTH, n_free / n_smp) : 1u;  // fewer free lanes than rays: the first n_free rays get one each
						uint32_t depth = 0;
						if (!alive) {
							const uint32_t q = (uint32_t)__popc(free_m & ((1u << lane) - 1u));
							const uint32_t k = q / h_per_ray;
							if (k < n_smp) { h_owner = (int)nth_set_bit(smp_m, k); depth = 1u + q % h_per_ray; }
						} else if (has_sample) {
							h_first = (uint32_t)__popc(smp_m & ((1u << lane) - 1u)) * h_per_ray;
							if (h_first >= n_free) h_per_ray = 0;
						}
						const int sl = h_owner < 0 ? (int)lane : h_owner;
						const V3 hro = v3(__shfl_sync(0xffffffffu, ro.x, sl), __shfl_sync(0xffffffffu, ro.y, sl), __shfl_sync(0xffffffffu, ro.z, sl));
						const V3 hrd = v3(__shfl_sync(0xffffffffu, rd.x, sl), __shfl_sync(0xffffffffu, rd.y, sl), __shfl_sync(0xffffffffu, rd.z, sl));
						float tj = __shfl_sync(0xffffffffu, t, sl);                  // the owner's t after its own sample
						const uint32_t hsteps = __shfl_sync(0xffffffffu, n_steps, sl);  // ... and its step count before it
						if (h_owner >= 0) {
							for (uint32_t i = 1; i < depth; ++i) tj = add(tj, calc_dt(tj, f.cone));  // the march of :671-688 over occupied samples
							pos = madd3(hrd, tj, hro);
							bool ok = hsteps + depth < MARCH_ITER - 1 && box_contains(f.rmin, f.rmax, pos);
							dt = calc_dt(tj, f.cone);
							if (ok && bitfield) {
								const uint32_t mip = (uint32_t)max(f.min_mip, mip_from_dt(dt, pos));
								ok = bitfield_at(cascaded_grid_idx_at(pos, mip), mip, bitfield);
							}
							has_sample = ok;
							if (ok) { dw = warp_direction(hrd); h_tnext = add(tj, dt); }
						}
					}
				}
			}
		}

		// ---- network inputs: generate_next_nerf_network_inputs :690 ----
		V3 pw = v3(0, 0, 0);
		float dtw = 0.0f;
		bool empty = false;
		Membrane mem;
		mem.op = -1; mem.tet = -1; mem.dob = 0.0f; mem.drb = 0.0f;
		bool need_old = false;
		V3 pw_old = pw;
		if (has_sample) {
			pw = warp_position(pos, f.tmin, f.tmax);
			dtw = warp_dt(dt);
			if (h_owner < 0) { ++n_steps; ++c_samples; }  // a helper's sample is counted by the owner when it composites it
			if (ops_on) {
				if (any_poisson) {  // membrane residuals are evaluated in deformed space (:2867-2883)
					poisson_one(ops, n_ops, pw, mem);
					need_old = mem.dob > 1e-9f && f.poisson_target;
				}
				pw_old = pw;
				map_one(ops, n_ops, pw, dw, empty);  // backward map into canonical space (:2896-2904)
			}
		}

		// ---- encode + fused MLPs. Pass 0 (rare) is the "old" density on the un-mapped position: the reference runs a second
		// full inference every round (:2892) but only the membrane's poisson-target blend ever reads it (:773), so it is
		// evaluated only in rounds where some lane needs it, density MLP only. Pass 1 is the real sample. One code copy.
		float sigma_old_raw = 0.0f;
		uint32_t dens[8], rgbo[8];
		int first_pass = 1;
		if (ops_on && any_poisson && f.poisson_target) first_pass = tc::tile_any(C, need_old) ? 0 : 1;
#ifdef NSB_PROFILE
		const long long c1 = clock64();
		long long enc_cycles = 0;
#endif
#pragma unroll 1
		for (int pass = first_pass; pass < 2; ++pass) {
			const bool old_pass = pass == 0;
#ifdef NSB_PROFILE
			const long long e0 = clock64();
#endif
			encode_to_a32(TB.a32, m, old_pass ? need_old : has_sample, old_pass ? pw_old : pw, tid);
#ifdef NSB_PROFILE
			enc_cycles += clock64() - e0;
#endif
			tc::run_network<ACC16>(C, phase, dw, old_pass, dens, rgbo);
			if (old_pass) {
				sigma_old_raw = h_lo(dens[0]);
				if (need_old) ++c_old;
			}
		}
#ifdef NSB_PROFILE
		const long long c2 = c1 + enc_cycles;
		const long long c3 = clock64();
#endif

		// ---- composite_kernel_nerf :750-955, split in two: what depends on the sample alone (any lane), then the ray's running sums (its owner lane) ----
		float s_alpha = 0.0f, s_e[3] = {0.0f, 0.0f, 0.0f}, s_z = 0.0f;
		float s_mask = 1.0f;  // glow_mode's mask_to_alpha: multiplies the sample's weight (:877-879)
		if (has_sample) {
			const uint32_t rtid = h_owner < 0 ? tid : (tid & ~31u) + (uint32_t)h_owner;  // the slot of the ray this sample belongs to
			const V3 cam_fwd = v3(f.cam1[6], f.cam1[7], f.cam1[8]);
			V3 cpos = unwarp_position(pw, f.tmin, f.tmax);
			float dtu = unwarp_dt(dtw);
			float sigma = network_to_density(h_lo(dens[0]), f.density_act);  // row 3 = density MLP out[0] (extract_density)
			float alpha;
			const bool membrane = OPS && mem.dob > 1e-9f;
			if (empty) {
				alpha = 0.0f;
			} else if (membrane) {
				float val;
				if (f.poisson_target) {
					float target = network_to_density(sigma_old_raw, f.density_act);
					val = fminf(fmaxf(target, sigma), sigma + mem.drb);
				} else {
					val = sigma + mem.drb;
				}
				alpha = 1.0f - __expf(-val * dtu);
			} else {
				alpha = 1.0f - __expf(-sigma * dtu);
			}
			if (fx_show_accel) alpha = 1.0f;  // show_accel >= 0: the occupancy cells themselves are drawn (:788-790)
			float rgb[3] = {network_to_rgb(h_lo(rgbo[0]), f.rgb_act), network_to_rgb(h_hi(rgbo[0]), f.rgb_act), network_to_rgb(h_lo(rgbo[1]), f.rgb_act)};
			if (fx_glow) {  // :807-903, before the mode overrides
				const float4 gl = glow_apply(fx_glow, f.glow_y_cutoff, cpos, v3(f.cam1[9], f.cam1[10], f.cam1[11]), rgb[0], rgb[1], rgb[2]);
				rgb[0] = gl.x; rgb[1] = gl.y; rgb[2] = gl.z; s_mask = gl.w;
			}
			if (f.mode != NSB_RENDER_SHADE) {
				const V3 ro = v3(TB.ray[R_OX][rtid], TB.ray[R_OY][rtid], TB.ray[R_OZ][rtid]);
				if (f.mode == NSB_RENDER_AO) { rgb[0] = rgb[1] = rgb[2] = alpha; }
				else if (f.mode == NSB_RENDER_POSITIONS && fx_show_accel) {  // one random colour per occupancy cell (:913-923)
					const uint32_t mip = (uint32_t)max(f.min_mip, mip_from_pos(cpos));
					const float res = (float)(GRIDSIZE >> mip);
					const int ix = (int)mul(cpos.x, res), iy = (int)mul(cpos.y, res), iz = (int)mul(cpos.z, res);
					Pcg32 rng;  // tcnn::default_rng_t rng(seed): pcg32::seed(initstate, initseq = 1)
					rng.state = 0; rng.inc = 3;
					rng.next_uint();
					rng.state += (uint64_t)(int64_t)(ix + iy * 232323 + iz * 727272);
					rng.next_uint();
					rgb[0] = fma_(-(float)mip, 0.25f, 1.0f);
					rgb[1] = rng.next_float();
					rgb[2] = rng.next_float();
				}
				else if (f.mode == NSB_RENDER_POSITIONS) { rgb[0] = (cpos.x - 0.5f) / 2.0f + 0.5f; rgb[1] = (cpos.y - 0.5f) / 2.0f + 0.5f; rgb[2] = (cpos.z - 0.5f) / 2.0f + 0.5f; }
				else if (f.mode == NSB_RENDER_DEPTH) { float z = dot3(cam_fwd, vsub(cpos, ro)) * f.depth_scale; rgb[0] = rgb[1] = rgb[2] = z; }
				else if (f.mode == NSB_RENDER_DISTANCE) { V3 q = vsub(cpos, ro); float z = sqrtf(dot3(q, q)) * f.depth_scale; rgb[0] = rgb[1] = rgb[2] = z; }
				else if (f.mode == NSB_RENDER_STEPSIZE) { float wdt = warp_dt(dtu); rgb[0] = rgb[1] = rgb[2] = wdt; }
			}
			if (membrane) {
				float alpha_N = 1.0f - __expf(-sigma * dtu);
				float alpha_R = 1.0f - __expf(-mem.dob * dtu);
				float w_N = alpha_N / (alpha_N + alpha_R), w_R = alpha_R / (alpha_N + alpha_R);
				float res[3];
				membrane_rgb(ops, mem, unwarp_direction(dw), res);
				rgb[0] = __fmaf_rn(w_R, res[0], w_N * rgb[0]);
				rgb[1] = __fmaf_rn(w_R, res[1], w_N * rgb[1]);
				rgb[2] = __fmaf_rn(w_R, res[2], w_N * rgb[2]);
			}
			s_alpha = alpha; s_e[0] = rgb[0]; s_e[1] = rgb[1]; s_e[2] = rgb[2];
			s_z = dot3(cam_fwd, vsub(cpos, v3(f.cam1[9], f.cam1[10], f.cam1[11])));
		}
		if constexpr (!OPS) {
			if (h_free) {  // warp-uniform: helpers post their samples, owners read them below
				if (h_owner >= 0) {
					mbox[M_TNEXT * MB] = h_tnext;
					if (has_sample) { mbox[M_ALPHA * MB] = s_alpha; mbox[M_E0 * MB] = s_e[0]; mbox[M_E1 * MB] = s_e[1]; mbox[M_E2 * MB] = s_e[2]; mbox[M_Z * MB] = s_z; }
				}
				__syncwarp();
			}
		}
		if (has_sample && h_owner < 0) {
			const float sat = 1.0f - f.min_T;  // rendering_min_transmittance test of composite_kernel_nerf :951
			float cr = TB.ray[R_CR][tid], cg = TB.ray[R_CG][tid], cb = TB.ray[R_CB][tid], ca = TB.ray[R_CA][tid];
			float ray_depth = TB.ray[R_DEPTH][tid], max_weight = TB.ray[R_MAXW][tid];
			auto accumulate = [&](const float alpha, const float e0, const float e1, const float e2, const float z) {
				float T = 1.0f - ca;
				float weight = alpha * T;
				cr = __fmaf_rn(e0, weight, cr);
				cg = __fmaf_rn(e1, weight, cg);
				cb = __fmaf_rn(e2, weight, cb);
				ca += weight;
				if (weight > max_weight) {
					max_weight = weight;
					ray_depth = z;
				}
				if (ca > sat) {
					float a = ca;
					cr = __fdiv_rn(cr, a); cg = __fdiv_rn(cg, a); cb = __fdiv_rn(cb, a); ca = __fdiv_rn(ca, a);
					finish(false, cr, cg, cb, ca, ray_depth);
				}
			};
			if (fx_glow & 4) {
				// mask_to_alpha: weight *= mask (a separate copy so that the unmasked expression — and its FMA contraction — stays exactly the reference's)
				float T = 1.0f - ca;
				float weight = s_alpha * T;
				weight *= s_mask;
				cr = __fmaf_rn(s_e[0], weight, cr);
				cg = __fmaf_rn(s_e[1], weight, cg);
				cb = __fmaf_rn(s_e[2], weight, cb);
				ca += weight;
				if (weight > max_weight) { max_weight = weight; ray_depth = s_z; }
				if (ca > sat) {
					float a = ca;
					cr = __fdiv_rn(cr, a); cg = __fdiv_rn(cg, a); cb = __fdiv_rn(cb, a); ca = __fdiv_rn(ca, a);
					finish(false, cr, cg, cb, ca, ray_depth);
				}
			} else {
				accumulate(s_alpha, s_e[0], s_e[1], s_e[2], s_z);
			}
			if constexpr (!OPS) {
				if (h_free && h_per_ray) {  // this ray's helpers: consecutive free lanes from free-rank h_first, in march order
					const float* wbox = mbox - lane;
					uint32_t hl = nth_set_bit(h_free, h_first);
					const uint32_t n_free = (uint32_t)__popc(h_free);
					for (uint32_t i = 0; alive && i < h_per_ray && h_first + i < n_free; ++i) {
						const float tn = wbox[M_TNEXT * MB + hl];
						if (tn < 0.0f) break;  // not an occupied in-bounds sample: the march resumes there next round
						++n_steps;
						++c_samples;
						accumulate(wbox[M_ALPHA * MB + hl], wbox[M_E0 * MB + hl], wbox[M_E1 * MB + hl], wbox[M_E2 * MB + hl], wbox[M_Z * MB + hl]);
						TB.ray[R_T][tid] = tn;
						hl = (uint32_t)__ffs((int)(h_free & ~((2u << hl) - 1u))) - 1u;  // next free lane
					}
				}
			}
			if (alive) {
				TB.ray[R_CR][tid] = cr; TB.ray[R_CG][tid] = cg; TB.ray[R_CB][tid] = cb; TB.ray[R_CA][tid] = ca;
				TB.ray[R_DEPTH][tid] = ray_depth; TB.ray[R_MAXW][tid] = max_weight;
			}
		}
#ifdef NSB_PROFILE
		cyc_acq += c1 - c0; cyc_enc += c2 - c1; cyc_mlp += c3 - c2; cyc_comp += clock64() - c3;
#endif
		++n_rounds;
	}

	tc::tc_fence_before();
	__syncthreads();  // all tiles done
	if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base, RENDER_TMEM_COLS);
	if (tid == 0) {
		atomicAdd(stats + ST_ROUNDS, (unsigned long long)n_rounds);
#ifdef NSB_PROFILE
		atomicAdd(stats + ST_CYC_ACQUIRE, (unsigned long long)cyc_acq);
		atomicAdd(stats + ST_CYC_ENCODE, (unsigned long long)cyc_enc);
		atomicAdd(stats + ST_CYC_MLP, (unsigned long long)cyc_mlp);
		atomicAdd(stats + ST_CYC_COMPOSITE, (unsigned long long)cyc_comp);
		atomicAdd(stats + ST_CYC_TOTAL, (unsigned long long)(clock64() - cyc_start));
#endif
		atomicAdd(stats + ST_CTAS, 1ull);
	}

	// counters: warp reduce, one atomic per warp
	unsigned long long c[3] = {c_hit, c_samples, c_old};  // per-thread counts fit 32 bits (one thread composites < 2^32 samples per frame)
	const int slot[3] = {ST_HIT, ST_SAMPLES, ST_OLD};
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		unsigned long long v = c[k];
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if (lane == 0 && v) atomicAdd(stats + slot[k], v);
	}
}

#include "nsb_render_ws.cuh"

// =====================================================================================================
// operator-level kernels (same device code as the fused renderer)
// =====================================================================================================
// NerfNetwork::inference_mixed_precision / ::density on a flat batch: one CTA = 128 samples.
template <bool DENSITY_ONLY, bool ACC16>
__global__ void __launch_bounds__(128) k_inference(const DevModel m, const float* __restrict__ coords, uint32_t n, __half* __restrict__ out,
                                                   uint32_t n_padded) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	tc::TileSmem& S = *reinterpret_cast<tc::TileSmem*>(smem_raw);
	const uint32_t tid = threadIdx.x;
	const uint32_t tmem_base = tc::tile_setup(S, m.w_image);
	uint32_t phase = 0;
	for (uint32_t base = blockIdx.x * 128u; base < n_padded; base += gridDim.x * 128u) {
		uint32_t i = base + tid;
		bool valid = i < n;
		V3 pw = v3(0, 0, 0), dw = v3(0.5f, 0.5f, 0.5f);
		if (valid) {
			pw = v3(coords[7 * (size_t)i], coords[7 * (size_t)i + 1], coords[7 * (size_t)i + 2]);
			dw = v3(coords[7 * (size_t)i + 4], coords[7 * (size_t)i + 5], coords[7 * (size_t)i + 6]);
		}
		uint32_t dens[8], rgbo[8];
		encode_to_a32(S.a32, m, valid, pw, tid);
		tc::run_network<ACC16>(tc::single_tile_ctx(S, tmem_base), phase, dw, DENSITY_ONLY, dens, rgbo);
		if (i < n_padded) {
			const uint32_t* src = DENSITY_ONLY ? dens : rgbo;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				out[(size_t)(2 * k) * n_padded + i] = __ushort_as_half((unsigned short)(src[k] & 0xffffu));
				out[(size_t)(2 * k + 1) * n_padded + i] = __ushort_as_half((unsigned short)(src[k] >> 16));
			}
			if (!DENSITY_ONLY) out[(size_t)3 * n_padded + i] = __ushort_as_half((unsigned short)(dens[0] & 0xffffu));  // extract_density
		}
	}
	tc::tile_teardown(S, tmem_base);
}

// tcnn GridEncoding inference alone: fp16 [32 x n_padded] row-major, same loop body as the fused kernel.
__global__ void __launch_bounds__(256) k_encode(const DevModel m, const float* __restrict__ coords, uint32_t n, __half* __restrict__ out, uint32_t n_padded) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float x = coords[7 * (size_t)i], y = coords[7 * (size_t)i + 1], z = coords[7 * (size_t)i + 2];
#pragma unroll 1
	for (uint32_t g = 0; g < MAX_LEVELS / 2; ++g) {
		__half2 v[2];
		const uint32_t mode = m.pair_mode[g];  // uniform
		if (mode == 1) encode_levels<2, 1>(m.levels + 2 * g, m.grid, x, y, z, v);
		else if (mode == 0) encode_levels<2, 0>(m.levels + 2 * g, m.grid, x, y, z, v);
		else encode_levels<2, 2>(m.levels + 2 * g, m.grid, x, y, z, v);
#pragma unroll
		for (uint32_t j = 0; j < 2; ++j) {
			out[(size_t)(4 * g + 2 * j) * n_padded + i] = __low2half(v[j]);
			out[(size_t)(4 * g + 2 * j + 1) * n_padded + i] = __high2half(v[j]);
		}
	}
}

// EditOperator::map_rays for operators [0, n_ops) of `ops` in reverse order; the mask is only ever SET (the caller clears it, testbed_nerf.cu:2898)
__global__ void k_map_rays(const DevOp* __restrict__ ops, int n_ops, float* __restrict__ coords, uint8_t* __restrict__ empty_mask, uint32_t n) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	V3 pw = v3(coords[7 * (size_t)i], coords[7 * (size_t)i + 1], coords[7 * (size_t)i + 2]);
	V3 dw = v3(coords[7 * (size_t)i + 4], coords[7 * (size_t)i + 5], coords[7 * (size_t)i + 6]);
	bool empty = false;
	map_one(ops, n_ops, pw, dw, empty);
	coords[7 * (size_t)i] = pw.x; coords[7 * (size_t)i + 1] = pw.y; coords[7 * (size_t)i + 2] = pw.z;
	coords[7 * (size_t)i + 4] = dw.x; coords[7 * (size_t)i + 5] = dw.y; coords[7 * (size_t)i + 6] = dw.z;
	if (empty) empty_mask[i] = 1;
}

// EditOperator::compute_poisson_full_residuals: outputs are written only where an operator's tet contains the sample (the caller clears, :2863-2866)
__global__ void k_poisson(const DevOp* __restrict__ ops, int n_ops, const float* __restrict__ coords, uint32_t n, float* __restrict__ sh,
                          float* __restrict__ od, float* __restrict__ rd) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Membrane mem;
	poisson_one(ops, n_ops, v3(coords[7 * (size_t)i], coords[7 * (size_t)i + 1], coords[7 * (size_t)i + 2]), mem);
	if (mem.op < 0) return;
	od[i] = mem.dob;
	rd[i] = mem.drb;
	for (int k = 0; k < 27; ++k) sh[27 * (size_t)i + k] = membrane_sh(ops, mem, k);
}

// EditOperator::map_positions: interpolate_tet_pos (cage_deformation.cu:136-192: no direction, and the vacated-region mask does NOT look at `copy`)
// and translate_in_box_pos (affine_duplication.cu:69-90), operators [0, n_ops) in reverse order; positions are `stride` floats apart
__global__ void k_map_positions(const DevOp* __restrict__ ops, int n_ops, float* __restrict__ pos, uint32_t stride, uint8_t* __restrict__ empty_mask, uint32_t n) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float* q = pos + (size_t)i * stride;
	V3 pw = v3(q[0], q[1], q[2]);
	bool empty = false;
	for (int o = n_ops - 1; o >= 0; --o) {
		const DevOp& op = ops[o];
		if (op.type == 0) {
			if (op.n_tets == 0) continue;
			bool in_deformed = false;
			if (box_contains(op.wbmin, op.wbmax, pw)) {
				V3 p = unwarp_position(pw, op.amin, op.amax);
				float b[4];
				int t = find_tet(op, p, b);
				if (t >= 0) {
					uint4 tv = __ldg(reinterpret_cast<const uint4*>(op.tets) + t);
					pw = warp_position(bary_mix(b, ldv(op.orig_verts, tv.x), ldv(op.orig_verts, tv.y), ldv(op.orig_verts, tv.z), ldv(op.orig_verts, tv.w)), op.amin, op.amax);
					in_deformed = true;
				}
			}
			if (!in_deformed && box_contains(op.owbmin, op.owbmax, pw)) {
				V3 p = unwarp_position(pw, op.amin, op.amax);
				int level = mip_from_pos(p);
				if (bitfield_at(cascaded_grid_idx_at(p, (uint32_t)level), (uint32_t)level, op.obits)) empty = true;
			}
		} else {
			V3 dw = v3(0.5f, 0.5f, 0.5f);
			DevOp tmp_free = op;  // translate_in_box_pos never touches a direction
			tmp_free.correct_dir = 0;
			affine_map(tmp_free, pw, dw, empty);
		}
	}
	q[0] = pw.x; q[1] = pw.y; q[2] = pw.z;
	if (empty && empty_mask) empty_mask[i] = 1;
}

// EditOperator::compute_poisson_residual_density (compute_poisson_residual_density_kernel, cage_deformation.cu:341-383):
// density_network_output[i] += (half) interpolated residual density, for the first tet that contains the (warped) position
__global__ void k_poisson_residual_density(const DevOp* __restrict__ ops, int n_ops, const float* __restrict__ pos, uint32_t stride, __half* __restrict__ density_out, uint32_t n) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* q = pos + (size_t)i * stride;
	const V3 pw = v3(q[0], q[1], q[2]);
	for (int o = n_ops - 1; o >= 0; --o) {
		const DevOp& op = ops[o];
		if (op.type != 0 || !op.apply_poisson || !op.has_poisson_data || op.n_tets == 0) continue;
		V3 p = unwarp_position(pw, op.amin, op.amax);
		if (!box_contains(op.bmin, op.bmax, p)) continue;
		float b[4];
		int t = find_tet(op, p, b);
		if (t < 0) continue;
		uint4 tv = __ldg(reinterpret_cast<const uint4*>(op.tets) + t);
		float res = bary_mix1(b, __ldg(op.rd + tv.x), __ldg(op.rd + tv.y), __ldg(op.rd + tv.z), __ldg(op.rd + tv.w));
		density_out[i] = __hadd(density_out[i], __float2half_rn(res));
	}
}

template <bool GENERAL>
__global__ void k_march_trace(const DevFrame f, const uint8_t* __restrict__ bitfield, const uint32_t* __restrict__ pixels, uint32_t n_pixels,
                              uint32_t max_samples, float* __restrict__ rec, uint32_t* __restrict__ idx, uint32_t* __restrict__ count) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pixels) return;
	uint32_t pix = pixels[i];
	uint32_t px = pix % (uint32_t)f.W, py = pix / (uint32_t)f.W;
	Ray r;
	uint32_t c = 0;
	if (GENERAL ? make_ray_general(f, px, py, r) : make_ray(f, px, py, r)) {
		V3 idir = v3(div_(1.0f, r.d.x), div_(1.0f, r.d.y), div_(1.0f, r.d.z));
		float t = fma_(ld_random_val(f.spp, pix * 786433u), calc_dt(r.t, f.cone), r.t);
		while (c < MARCH_ITER) {
			float dt; V3 pos; uint32_t mip, cell;
			if (!next_occupied(f, bitfield, r.o, r.d, idir, t, dt, pos, mip, cell)) break;
			if (c < max_samples) {
				float* rr = rec + ((size_t)i * max_samples + c) * 5;
				rr[0] = t; rr[1] = dt; rr[2] = pos.x; rr[3] = pos.y; rr[4] = pos.z;
				uint32_t* ii = idx + ((size_t)i * max_samples + c) * 2;
				ii[0] = mip; ii[1] = cell;
			}
			++c;
			t = add(t, dt);
		}
	}
	count[i] = c;
}

// ---- density grid -> occupancy bitfield (testbed_nerf.cu:514-555, :3642-3658) -----------------------------------------
__global__ void k_grid_mean_partial(const float* __restrict__ grid, uint32_t n, double* __restrict__ partial) {
	// fixed-shape tree: 256 threads x 64 values each per block, pairwise in shared memory (deterministic order)
	__shared__ double sh[256];
	uint32_t base = blockIdx.x * 256u * 64u + threadIdx.x * 64u;
	double acc = 0.0;
	for (uint32_t k = 0; k < 64; ++k) { uint32_t i = base + k; if (i < n) acc += (double)fmaxf(grid[i], 0.0f); }
	sh[threadIdx.x] = acc;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
	if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249u;
	x = (x | (x >> 2)) & 0xc30c30c3u;
	x = (x | (x >> 4)) & 0x0f00f00fu;
	x = (x | (x >> 8)) & 0xff0000ffu;
	x = (x | (x >> 16)) & 0x0000ffffu;
	return x;
}
__global__ void k_bitfield_max_pool(uint32_t n, const uint8_t* __restrict__ prev_level, uint8_t* __restrict__ next_level) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint8_t b = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) b |= prev_level[(size_t)i * 8 + j] > 0 ? (uint8_t)(1u << j) : (uint8_t)0;
	uint32_t x = morton3D_invert(i >> 0) + GRIDSIZE / 8, y = morton3D_invert(i >> 1) + GRIDSIZE / 8, z = morton3D_invert(i >> 2) + GRIDSIZE / 8;
	next_level[morton3D(x, y, z)] |= b;
}

// ---- occupancy update through the edit operators (testbed_nerf.cu:3533-3639) ------------------------------------------
// One fused pass per grid sample: draw (common_nerf.cu:179-208) -> map_positions in reverse operator order (:3593-3599)
// -> hash encode + density MLP on tcgen05 (NerfNetwork::density) -> activate (:3522-3531) -> membrane residual density
// (cage_deformation.cu:341-383) -> max-splat (:447-463). The reference materialises positions, indices, a bool mask and the
// 16 x N fp16 density matrix between seven launches; here a sample lives in registers/shared memory/TMEM from draw to splat.
struct GridUpdateArgs {
	uint32_t n_uniform, n_total, n_cascades, step;
	uint64_t rng_state, rng_inc;
	float amin[3], amax[3];
	int density_activation, apply_ops;
};
template <bool ACC16>
__global__ void __launch_bounds__(128) k_density_grid_update(const DevModel m, const DevOp* __restrict__ ops, int n_ops, const GridUpdateArgs a,
                                                             const float* __restrict__ grid, float* __restrict__ grid_tmp) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	tc::TileSmem& S = *reinterpret_cast<tc::TileSmem*>(smem_raw);
	const uint32_t tid = threadIdx.x;
	const uint32_t tmem_base = tc::tile_setup(S, m.w_image);
	uint32_t phase = 0;
	const uint32_t n_padded = ((a.n_total + 127u) / 128u) * 128u;
	for (uint32_t base = blockIdx.x * 128u; base < n_padded; base += gridDim.x * 128u) {
		const uint32_t i = base + tid;
		const bool valid = i < a.n_total;
		V3 pw = v3(0, 0, 0), dw = v3(0.5f, 0.5f, 0.5f);
		uint32_t idx = 0;
		if (valid) {
			// the reference runs two launches (uniform: any trained cell; non-uniform: occupied cells), m_rng advanced by 2^32 between
			const bool second = i >= a.n_uniform;
			uint32_t li = second ? i - a.n_uniform : i;
			const uint32_t n_el = second ? a.n_total - a.n_uniform : a.n_uniform;
			const float thresh = second ? 0.01f : -0.01f;
			// Work assignment (not semantics): the reference's thread li draws cell (li*A + B) mod 128^3, a pseudo-random permutation,
			// so a warp's 32 samples scatter over the whole volume. Every sample is independent and the splat is an order-free max, so
			// thread t instead takes the sample whose first-choice cell is t mod 128^3: a warp covers a 4x4x2 Morton block of cells and
			// its hash-grid gathers share sectors (53369 = 56924617^-1 mod 2^21). Exact for whole multiples of 128^3 samples (the reference's post-edit setting).
			if (li < (n_el & ~(GRIDVOL - 1u)))
				li = (li & ~(GRIDVOL - 1u)) | ((((li & (GRIDVOL - 1u)) - 96925573u) * 53369u - a.step * n_el) & (GRIDVOL - 1u));
			Pcg32 rng;
			rng.state = a.rng_state; rng.inc = a.rng_inc;
			if (second) rng.advance(1ull << 32);
			rng.advance((uint64_t)(li * 4u));
			const uint32_t level = (uint32_t)mul(rng.next_float(), (float)a.n_cascades) % a.n_cascades;
			for (uint32_t j = 0; j < 10; ++j) {
				idx = ((li + a.step * n_el) * 56924617u + j * 19349663u + 96925573u) % GRIDVOL;
				idx += level * GRIDVOL;
				if (grid[idx] > thresh) break;
			}
			const uint32_t pos_idx = idx % GRIDVOL;
			const float x = (float)morton3D_invert(pos_idx >> 0), y = (float)morton3D_invert(pos_idx >> 1), z = (float)morton3D_invert(pos_idx >> 2);
			const float rx = rng.next_float(), ry = rng.next_float(), rz = rng.next_float();
			const float sc = __uint_as_float((127u + level) << 23);  // scalbnf(1, level)
			V3 p = v3(fma_(sub(div_(add(x, rx), (float)GRIDSIZE), 0.5f), sc, 0.5f), fma_(sub(div_(add(y, ry), (float)GRIDSIZE), 0.5f), sc, 0.5f),
			          fma_(sub(div_(add(z, rz), (float)GRIDSIZE), 0.5f), sc, 0.5f));
			pw = warp_position(p, a.amin, a.amax);
			if (a.apply_ops && n_ops > 0) {
				bool empty = false;  // the reference fills empty_mask here but its consumer clear_empty_space is a no-op (:2759-2770)
				V3 d = dw;
				map_one(ops, n_ops, pw, d, empty);
			}
		}
		uint32_t dens[8], rgbo[8];
		encode_to_a32(S.a32, m, valid, pw, tid);
		tc::run_network<ACC16>(tc::single_tile_ctx(S, tmem_base), phase, dw, true, dens, rgbo);
		if (valid) {
			__half h = __float2half_rn(network_to_density(__half2float(__ushort_as_half((unsigned short)(dens[0] & 0xffffu))), a.density_activation));
			if (a.apply_ops) {
				for (int o = n_ops - 1; o >= 0; --o) {  // :3612-3620; looked up at the MAPPED position, as the reference does
					const DevOp& op = ops[o];
					if (op.type != 0 || !op.apply_poisson || !op.has_poisson_data || op.n_tets == 0) continue;
					V3 p = unwarp_position(pw, op.amin, op.amax);
					if (!box_contains(op.bmin, op.bmax, p)) continue;
					float b[4];
					int t = find_tet(op, p, b);
					if (t < 0) continue;
					uint4 tv = __ldg(reinterpret_cast<const uint4*>(op.tets) + t);
					float res = bary_mix1(b, __ldg(op.rd + tv.x), __ldg(op.rd + tv.y), __ldg(op.rd + tv.z), __ldg(op.rd + tv.w));
					h = __hadd(h, __float2half_rn(res));
				}
			}
			const float optical_thickness = mul(__half2float(h), MIN_STEP());
			atomicMax(reinterpret_cast<unsigned int*>(grid_tmp) + idx, __float_as_uint(optical_thickness));  // uint order == float order for >= 0
		}
	}
	tc::tile_teardown(S, tmem_base);
}
__global__ void k_ema_grid(uint32_t n, float decay, float* __restrict__ grid, const float* __restrict__ grid_tmp) {  // :483-506
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float prev = grid[i];
	grid[i] = prev < 0.0f ? prev : fmaxf(mul(prev, decay), grid_tmp[i]);
}
__global__ void k_grid_mean_final(const double* __restrict__ partial, uint32_t n_blocks, uint32_t n0, float* __restrict__ thresh_out) {
	double sum = 0.0;
	for (uint32_t k = 0; k < n_blocks; ++k) sum += partial[k];  // fixed order
	const float mean = (float)(sum / (double)n0);
	thresh_out[0] = mean < 0.01f ? mean : 0.01f;  // std::min(NERF_MIN_OPTICAL_THICKNESS(), mean) (:524)
	thresh_out[1] = mean;
}
__global__ void k_grid_to_bitfield_dev(uint32_t n_bytes, const float* __restrict__ grid, uint8_t* __restrict__ bits, const float* __restrict__ thresh_ptr) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_bytes) return;
	const float thresh = thresh_ptr[0];
	uint8_t b = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) b |= grid[(size_t)i * 8 + j] > thresh ? (uint8_t)(1u << j) : (uint8_t)0;
	bits[i] = b;
}

// ---- frame post-process: accumulate_kernel (render_buffer.cu:217-258) and tonemap_kernel (:471-499) -------------------
__device__ __forceinline__ float linear_to_srgb(float l) {  // common_device.cuh:53-59
	if (l < 0.0031308f) return 12.92f * l;
	return 1.055f * powf(l, 0.41666f) - 0.055f;
}
__global__ void k_accumulate(int n, const float4* __restrict__ frame, float4* __restrict__ acc, float sample_count, int color_space) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 c = frame[i], t = acc[i];
	const float inv = sample_count + 1.0f;
	if (color_space == NSB_COLOR_VISPOSNEG) {
		float val = c.x - c.y, tmp_val = t.x - t.y;
		tmp_val = (tmp_val * sample_count + val) / inv;
		t.x = fmaxf(tmp_val, 0.0f);
		t.y = fmaxf(-tmp_val, 0.0f);
	} else {
		if (color_space == NSB_COLOR_SRGB) { c.x = linear_to_srgb(c.x); c.y = linear_to_srgb(c.y); c.z = linear_to_srgb(c.z); }
		t.x = (t.x * sample_count + c.x) / inv;
		t.y = (t.y * sample_count + c.y) / inv;
		t.z = (t.z * sample_count + c.z) / inv;
	}
	t.w = (t.w * sample_count + c.w) / inv;
	acc[i] = t;
}
__device__ __forceinline__ void tonemap_curve(float* x, int curve) {  // render_buffer.cu:262-310
	if (curve == NSB_TONEMAP_IDENTITY) return;
	for (int k = 0; k < 3; ++k) x[k] = fmaxf(x[k], 0.0f);
	float k0, k1, k2, k3, k4, k5;
	if (curve == NSB_TONEMAP_ACES) {
		k0 = 0.6f * 0.6f * 2.51f; k1 = 0.6f * 0.03f; k2 = 0.0f; k3 = 0.6f * 0.6f * 2.43f; k4 = 0.6f * 0.59f; k5 = 0.14f;
	} else if (curve == NSB_TONEMAP_HABLE) {
		const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
		k0 = A * F - A * E; k1 = C * B * F - B * E; k2 = 0.0f; k3 = A * F; k4 = B * F; k5 = D * F * F;
		const float W = 11.2f;
		const float nom = k0 * (W * W) + k1 * W + k2, denom = k3 * (W * W) + k4 * W + k5;
		const float white_scale = denom / nom;
		k0 = 4.0f * k0 * white_scale; k1 = 2.0f * k1 * white_scale; k2 = k2 * white_scale; k3 = 4.0f * k3; k4 = 2.0f * k4;
	} else {  // Reinhard
		float Y = 0.2126f * x[0] + 0.7152f * x[1] + 0.0722f * x[2];
		float s = 1.0f / (Y + 1.0f);
		for (int k = 0; k < 3; ++k) x[k] *= s;
		return;
	}
	for (int k = 0; k < 3; ++k) {
		float sq = x[k] * x[k];
		x[k] = (sq * k0 + k1 * x[k] + k2) / (k3 * sq + k4 * x[k] + k5);
	}
}
__global__ void k_tonemap(int n, const float4* __restrict__ acc, float4* __restrict__ out, NsbTonemap p) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float bg[4] = {p.background_color[0], p.background_color[1], p.background_color[2], p.background_color[3]};
	if (p.color_space != NSB_COLOR_SRGB) for (int k = 0; k < 3; ++k) bg[k] = srgb_to_linear(bg[k]);
	float4 c4 = acc[i];
	float c[3] = {c4.x, c4.y, c4.z};
	float weight = (1.0f - c4.w) * bg[3];
	for (int k = 0; k < 3; ++k) c[k] += bg[k] * weight;
	float a = c4.w + weight;
	if (p.color_space == NSB_COLOR_SRGB) for (int k = 0; k < 3; ++k) c[k] = srgb_to_linear(c[k]);
	const float e = powf(2.0f, p.exposure);
	for (int k = 0; k < 3; ++k) c[k] *= e;
	tonemap_curve(c, p.tonemap_curve);
	if (p.output_color_space == NSB_COLOR_SRGB) for (int k = 0; k < 3; ++k) c[k] = linear_to_srgb(c[k]);
	if (p.clamp_output_color) { for (int k = 0; k < 3; ++k) c[k] = fminf(fmaxf(c[k], 0.0f), 1.0f); a = fminf(fmaxf(a, 0.0f), 1.0f); }
	out[i] = make_float4(c[0], c[1], c[2], a);
}

// packed tile buffers for the single multi-GPU framebuffer gather
__global__ void k_pack_tiles(const float4* __restrict__ fb, const float* __restrict__ depth, int W, int H, int tiles_x, int n_tiles, int rank, int world,
                             float4* __restrict__ dst, float* __restrict__ dst_depth, int unpack) {
	uint32_t local = blockIdx.x;
	uint32_t tile = (uint32_t)rank + local * (uint32_t)world;
	if (tile >= (uint32_t)n_tiles) return;
	uint32_t lane = threadIdx.x;  // 0..127
	uint32_t px = (tile % (uint32_t)tiles_x) * TILE_W + (lane % TILE_W), py = (tile / (uint32_t)tiles_x) * TILE_H + (lane / TILE_W);
	if (px >= (uint32_t)W || py >= (uint32_t)H) return;
	size_t p = (size_t)px + (size_t)W * py, q = (size_t)local * TILE_PIXELS + lane;
	if (!unpack) {
		dst[q] = fb[p];
		if (dst_depth && depth) dst_depth[q] = depth[p];
	} else {
		const_cast<float4*>(fb)[p] = dst[q];
		if (dst_depth && depth) const_cast<float*>(depth)[p] = dst_depth[q];
	}
}

// all the OTHER ranks' shards of one all-gather -> the framebuffer (and depth), ONE launch: block = (local tile, source rank)
__global__ void k_unpack_gathered(const float4* __restrict__ gathered, const float* __restrict__ gathered_depth, int W, int H, int tiles_x, int n_tiles, int skip_rank,
                                  int world, uint32_t rank_stride /*floats between two ranks' shards, in BOTH arrays*/, float4* __restrict__ fb, float* __restrict__ depth) {
	const uint32_t local = blockIdx.x, rank = blockIdx.y;
	if ((int)rank == skip_rank) return;
	const uint32_t tile = rank + local * (uint32_t)world;
	if (tile >= (uint32_t)n_tiles) return;
	const uint32_t lane = threadIdx.x;
	const uint32_t px = (tile % (uint32_t)tiles_x) * TILE_W + (lane % TILE_W), py = (tile / (uint32_t)tiles_x) * TILE_H + (lane / TILE_W);
	if (px >= (uint32_t)W || py >= (uint32_t)H) return;
	const size_t p = (size_t)px + (size_t)W * py, q = (size_t)local * TILE_PIXELS + lane;
	fb[p] = gathered[(size_t)rank * (rank_stride / 4) + q];
	if (depth && gathered_depth) depth[p] = gathered_depth[(size_t)rank * rank_stride + q];
}

// =====================================================================================================
// host side: context, uploads, C ABI
// =====================================================================================================
#include "nsb_edit_rebuild.cuh"
static thread_local std::string g_last_error;
static NsbStatus fail(NsbStatus s, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_last_error = buf;
	return s;
}
#define CU(call)                                                                                                 \
	do {                                                                                                         \
		cudaError_t e__ = (call);                                                                                \
		if (e__ != cudaSuccess) return fail(NSB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

struct NsbContext {
	int device = 0;
	int sm_count = 0;
	int ctas_per_sm = 1;
	int inference_ctas_per_sm = 1;
	int grid_update_ctas_per_sm = 1;
	bool has_model = false, has_occ = false;
	NsbModelDesc desc{};
	DevModel model{};
	__half2* d_grid = nullptr;
	uint8_t* d_wimage = nullptr;
	__half* d_wrow = nullptr;
	uint8_t* d_bitfield = nullptr;
	float* d_density_grid = nullptr;   // running density grid, 5 x 128^3 (Testbed::Nerf::density_grid, testbed.h:622)
	float* d_density_tmp = nullptr;    // density_grid_tmp scratch of an update
	double* d_mean_partial = nullptr;
	float* d_thresh = nullptr;         // [0] bitfield threshold, [1] mean of cascade 0
	DevOp* d_ops = nullptr;
	int n_ops = 0;
	int any_poisson = 0;
	std::vector<void*> op_allocs;
	std::vector<DevOp> h_ops;          // host copy of d_ops (pointers are device pointers)
	struct CageRebuild {               // device scratch of nsb_cage_deform, one per operator slot
		float* d_mvc = nullptr; uint32_t n_cv = 0, n_vertices = 0;
		float* d_cage = nullptr;
		uint32_t* d_counts = nullptr; uint32_t* d_block_sums = nullptr; uint32_t* d_total = nullptr;
		nsb::rebuild::Mark* d_marks = nullptr; uint32_t marks_cap = 0; unsigned int* d_n_marks = nullptr;
		uint32_t* d_idx = nullptr; uint64_t idx_cap = 0;   // owned replacement of the uploaded tet_lut_idx once it had to grow
		float* d_boxes = nullptr;      // [12] bbox min/max, warped min/max; [12..17] scene aabb
		float* d_gamma = nullptr; float* d_mem_small = nullptr; uint32_t mem_n_cv = 0;  // nsb_cage_set_membrane scratch
		uint64_t n_idx = 0;
	};
	std::vector<CageRebuild> rb;
	uint32_t* d_counters = nullptr;  // [0] rays queued by k_prepare_rays, [1] fetch cursor of k_render_fused
	unsigned long long* d_stats = nullptr;
	RayRec* d_list = nullptr;
	size_t list_capacity = 0;
	float* d_ray_od = nullptr;     // origin + direction per queue slot, general-camera frames only (allocated on the first such frame)
	size_t ray_od_capacity = 0;
	int acc16 = 1;                 // MLP accumulator policy: 1 fp16 TMEM accumulators (default: the reference's wmma __half fragments), 0 fp32 (nsb_set_mlp_accumulator)
	int use_ws = 0;                // NSB_WS=1: frames without operators go through the warp-specialised kernel (experimental; default k_render_fused<false>)
	int helpers = 0;               // helper lanes of k_render_fused (experiment build only, -DNSB_X_HELP=1; measured slower: profiles/README.md): 0 off (default), 1 once the queue is exhausted, 2 always (NSB_HELPERS)
	int refill_thr = 2;            // lanes of a warp refill when at most this many of its rays are alive (31: immediately)
	int dda_budget = DDA_BUDGET;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;  // start, end, between k_prepare_rays and k_render_fused
	bool timed = false;
	uint32_t launches = 0;
	float* d_fb = nullptr;
	float* d_depth = nullptr;
	size_t fb_pixels = 0;
	cudaStream_t stream = nullptr;
	// nsb_render_host_async: two framebuffers in flight, the D2H copy of one overlaps the render of the other
	float* d_fb2[2] = {nullptr, nullptr};
	float* d_depth2[2] = {nullptr, nullptr};
	size_t fb2_pixels = 0;
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t render_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
	uint64_t host_seq = 0;  // tickets issued
};

static uint32_t next_multiple_u32(uint32_t v, uint32_t m) { return ((v + m - 1) / m) * m; }

static bool build_levels(const NsbModelDesc* d, DevLevel* L, uint64_t* n_grid_entries) {
	if (!d || d->n_levels != 16 || d->n_features_per_level != 2 || d->n_neurons != 64 || d->n_hidden_density != 1 || d->n_hidden_rgb != 2 ||
	    d->sh_degree != 4 || d->log2_hashmap_size == 0 || d->log2_hashmap_size > 24 || d->base_resolution == 0 || !(d->per_level_scale > 0.0f))
		return false;
	float l2s = log2f(d->per_level_scale);
	uint64_t offset = 0;
	for (uint32_t l = 0; l < d->n_levels; ++l) {
		float scale = exp2f((float)l * l2s) * (float)d->base_resolution - 1.0f;  // tcnn grid_scale
		uint32_t res = (uint32_t)ceilf(scale) + 1u;                                // tcnn grid_resolution
		double dense = (double)res * res * res;
		uint32_t max_params = 0xffffffffu / 2;
		uint32_t n = dense > (double)max_params ? max_params : (uint32_t)dense;
		n = next_multiple_u32(n, 8u);
		n = n < (1u << d->log2_hashmap_size) ? n : (1u << d->log2_hashmap_size);
		L[l].scale = scale;
		L[l].res = res;
		L[l].base = reinterpret_cast<const __half2*>((uintptr_t)offset * sizeof(__half2));  // entry offset; rebased onto the table by nsb_upload_model
		L[l].size = n;
		L[l].hashed = dense > (double)n ? 1u : 0u;
		L[l].res2 = L[l].hashed ? 0u : res * res;
		L[l].mask = (n & (n - 1)) == 0 ? n - 1 : 0u;
		offset += n;
	}
	*n_grid_entries = offset;
	return true;
}
static const uint64_t kMlpParams = 64 * 32 + 16 * 64 + 64 * 32 + 64 * 64 + 16 * 64;

extern "C" int nsb_abi_version(void) { return NSB_ABI_VERSION; }
extern "C" const char* nsb_last_error(void) { return g_last_error.c_str(); }

extern "C" NsbStatus nsb_model_n_params(const NsbModelDesc* desc, uint64_t* n_params) {
	DevLevel L[MAX_LEVELS];
	uint64_t n = 0;
	if (!n_params || !build_levels(desc, L, &n)) return fail(NSB_ERR_INVALID, "unsupported model description (only the configs/nerf/base.json topology is implemented)");
	*n_params = kMlpParams + 2 * n;
	return NSB_OK;
}

extern "C" NsbStatus nsb_create(int device, NsbContext** out) {
	if (!out) return fail(NSB_ERR_INVALID, "out is null");
	int n_dev = 0;
	CU(cudaGetDeviceCount(&n_dev));
	if (n_dev == 0) return fail(NSB_ERR_CUDA, "no CUDA device: nerfshop_b200 has no CPU fallback");
	if (device < 0) CU(cudaGetDevice(&device));
	if (device >= n_dev) return fail(NSB_ERR_INVALID, "device %d out of range (%d devices)", device, n_dev);
	CU(cudaSetDevice(device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) return fail(NSB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
	NsbContext* c = new NsbContext();
	c->device = device;
	c->sm_count = prop.multiProcessorCount;
	CU(cudaMalloc(&c->d_counters, 2 * sizeof(uint32_t)));
	CU(cudaMalloc(&c->d_stats, ST_N * sizeof(unsigned long long)));
	CU(cudaMemset(c->d_stats, 0, ST_N * sizeof(unsigned long long)));
	CU(cudaEventCreate(&c->ev0));
	CU(cudaEventCreate(&c->ev1));
	CU(cudaEventCreate(&c->evm));
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; ++i) {
		CU(cudaEventCreateWithFlags(&c->render_done[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&c->copy_done[i], cudaEventDisableTiming));
	}
	// Shared memory and L1 share one 256 KB array per SM: every kernel asks for exactly the carve-out its resident CTAs need and
	// leaves the rest to L1, which serves the hash-grid gathers (the default heuristic picks a carve-out too small for even one
	// CTA; "max shared" starves L1 — the tile loop runs 44 % faster with 92 KB of L1 than with 28 KB, profiles/README.md).
	auto set_smem = [&](const void* func, size_t bytes_per_cta, int ctas) -> cudaError_t {
		cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_per_cta);
		if (e != cudaSuccess) return e;
		const size_t need = (bytes_per_cta + 1024) * (size_t)ctas;  // 1 KB per CTA is reserved by the system
		int pct = (int)((need * 100 + prop.sharedMemPerMultiprocessor - 1) / prop.sharedMemPerMultiprocessor);
		if (pct > 100) pct = 100;
		return cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
	};
	// k_render_fused: ONE CTA of NSB_TILES (NSB_TILES_OPS with operators) tiles per SM (768 x 80 / 640 x 96 registers = the register file; TMEM 64 columns
	// per tile; shared memory 20 KB weights + 23 KB per tile).
	c->ctas_per_sm = 1;
	CU(set_smem((const void*)k_render_fused<false, false>, sizeof(RenderSmemT<render_tiles(false)>), 1));
	CU(set_smem((const void*)k_render_fused<true, false>, sizeof(RenderSmemT<render_tiles(true)>), 1));
	CU(set_smem((const void*)k_render_fused<false, true>, sizeof(RenderSmemT<render_tiles(false)>), 1));
	CU(set_smem((const void*)k_render_fused<true, true>, sizeof(RenderSmemT<render_tiles(true)>), 1));
	cudaFuncAttributes fa;
	CU(cudaFuncGetAttributes(&fa, k_render_fused<false, false>));
	const int by_smem_tile = (int)(prop.sharedMemPerMultiprocessor / (sizeof(tc::TileSmem) + 1024));
	{   // the operator-level kernels: one tile per CTA, residency from registers / shared memory
		cudaFuncAttributes fi;
		CU(cudaFuncGetAttributes(&fi, k_inference<false, false>));
		int r = fi.numRegs > 0 ? (int)(prop.regsPerMultiprocessor / (fi.numRegs * 128)) : 4;
		c->inference_ctas_per_sm = r < by_smem_tile ? r : by_smem_tile;
		if (c->inference_ctas_per_sm > 5) c->inference_ctas_per_sm = 5;
		if (c->inference_ctas_per_sm < 1) c->inference_ctas_per_sm = 1;
		CU(cudaFuncGetAttributes(&fi, k_density_grid_update<false>));
		r = fi.numRegs > 0 ? (int)(prop.regsPerMultiprocessor / (fi.numRegs * 128)) : 4;
		c->grid_update_ctas_per_sm = r < by_smem_tile ? r : by_smem_tile;
		if (c->grid_update_ctas_per_sm > 5) c->grid_update_ctas_per_sm = 5;
		if (c->grid_update_ctas_per_sm < 1) c->grid_update_ctas_per_sm = 1;
		CU(set_smem((const void*)k_inference<false, false>, sizeof(tc::TileSmem), c->inference_ctas_per_sm));
		CU(set_smem((const void*)k_inference<true, false>, sizeof(tc::TileSmem), c->inference_ctas_per_sm));
		CU(set_smem((const void*)k_inference<false, true>, sizeof(tc::TileSmem), c->inference_ctas_per_sm));
		CU(set_smem((const void*)k_inference<true, true>, sizeof(tc::TileSmem), c->inference_ctas_per_sm));
		CU(set_smem((const void*)k_density_grid_update<false>, sizeof(tc::TileSmem), c->grid_update_ctas_per_sm));
		CU(set_smem((const void*)k_density_grid_update<true>, sizeof(tc::TileSmem), c->grid_update_ctas_per_sm));
	}
	if (const char* e = getenv("NSB_CARVEOUT")) {  // experiments
		CU(cudaFuncSetAttribute(k_render_fused<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
		CU(cudaFuncSetAttribute(k_render_fused<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
	}
	CU(set_smem((const void*)k_render_ws, sizeof(ws::Smem), 1));
	if (const char* e = getenv("NSB_WS")) c->use_ws = atoi(e) != 0;
	if (const char* e = getenv("NSB_MLP_ACC16")) c->acc16 = atoi(e) != 0;
	if (const char* e = getenv("NSB_REFILL_THR")) { int v = atoi(e); if (v >= 0 && v <= 31) c->refill_thr = v; }
	if (const char* e = getenv("NSB_HELPERS")) { int v = atoi(e); if (v >= 0 && v <= 2) c->helpers = v; }
	if (const char* e = getenv("NSB_CHUNK")) { int v = atoi(e); if (v >= 0 && v <= 65535) c->refill_thr |= v << 8; }
	if (const char* e = getenv("NSB_DDA_BUDGET")) { int v = atoi(e); if (v >= 0 && v <= 1024) c->dda_budget = v; }
	if (getenv("NSB_VERBOSE"))
		fprintf(stderr, "[nsb] device %d: %d SMs; k_render_fused %d tiles/CTA (%d with operators), %d regs, %zu B smem; inference %d CTAs/SM, grid update %d CTAs/SM\n", device, c->sm_count,
		        render_tiles(false), render_tiles(true), fa.numRegs, sizeof(RenderSmemT<render_tiles(false)>), c->inference_ctas_per_sm, c->grid_update_ctas_per_sm);
	*out = c;
	return NSB_OK;
}

static void free_ops(NsbContext* c) {
	for (void* p : c->op_allocs) cudaFree(p);
	c->op_allocs.clear();
	for (auto& r : c->rb) {
		cudaFree(r.d_mvc); cudaFree(r.d_cage); cudaFree(r.d_counts); cudaFree(r.d_block_sums); cudaFree(r.d_total); cudaFree(r.d_marks);
		cudaFree(r.d_n_marks); cudaFree(r.d_idx); cudaFree(r.d_boxes); cudaFree(r.d_gamma); cudaFree(r.d_mem_small);
	}
	c->rb.clear();
	c->h_ops.clear();
	if (c->d_ops) cudaFree(c->d_ops);
	c->d_ops = nullptr;
	c->n_ops = 0;
	c->any_poisson = 0;
}

extern "C" NsbStatus nsb_destroy(NsbContext* c) {
	if (!c) return NSB_OK;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	free_ops(c);
	cudaFree(c->d_grid); cudaFree(c->d_wimage); cudaFree(c->d_wrow); cudaFree(c->d_bitfield);
	cudaFree(c->d_density_grid); cudaFree(c->d_density_tmp); cudaFree(c->d_mean_partial); cudaFree(c->d_thresh);
	cudaFree(c->d_counters); cudaFree(c->d_stats); cudaFree(c->d_fb); cudaFree(c->d_depth); cudaFree(c->d_list); cudaFree(c->d_ray_od);
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	if (c->evm) cudaEventDestroy(c->evm);
	if (c->stream) cudaStreamDestroy(c->stream);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	for (int i = 0; i < 2; ++i) {
		if (c->render_done[i]) cudaEventDestroy(c->render_done[i]);
		if (c->copy_done[i]) cudaEventDestroy(c->copy_done[i]);
		cudaFree(c->d_fb2[i]); cudaFree(c->d_depth2[i]);
	}
	delete c;
	return NSB_OK;
}

// [N x K] row-major fp16 -> UMMA K-major no-swizzle operand image (nsb_tc.cuh header comment)
static void to_operand_layout(const uint16_t* W, uint32_t N, uint32_t K, uint8_t* dst) {
	for (uint32_t n = 0; n < N; ++n)
		for (uint32_t k = 0; k < K; ++k) {
			size_t off = (size_t)(k / 8) * (N * 16) + (size_t)n * 16 + (k % 8) * 2;
			memcpy(dst + off, W + (size_t)n * K + k, 2);
		}
}

// params_on_device: the block is read where tcnn::Trainer keeps it; only the 20 KB of MLP weights pass through the host (to be laid out as UMMA operands)
static NsbStatus upload_model_impl(NsbContext* c, const NsbModelDesc* desc, const uint16_t* params, uint64_t n_params, bool params_on_device) {
	if (!c || !desc || !params) return fail(NSB_ERR_INVALID, "null argument");
	CU(cudaSetDevice(c->device));
	DevModel m{};
	uint64_t n_grid = 0;
	if (!build_levels(desc, m.levels, &n_grid)) return fail(NSB_ERR_INVALID, "unsupported model description (only the configs/nerf/base.json topology is implemented)");
	if (n_params != kMlpParams + 2 * n_grid) return fail(NSB_ERR_INVALID, "n_params %llu != expected %llu", (unsigned long long)n_params, (unsigned long long)(kMlpParams + 2 * n_grid));
	CU(cudaDeviceSynchronize());
	cudaFree(c->d_grid); cudaFree(c->d_wimage); cudaFree(c->d_wrow);
	c->d_grid = nullptr; c->d_wimage = nullptr; c->d_wrow = nullptr;
	// block order: density MLP, rgb MLP, hash grid (nerf_network_full.h:316-349)
	std::vector<uint8_t> image(tc::W_BYTES);
	std::vector<uint16_t> mlp_host;
	if (params_on_device) {
		mlp_host.resize(kMlpParams);
		CU(cudaMemcpy(mlp_host.data(), params, kMlpParams * 2, cudaMemcpyDeviceToHost));
	}
	const uint16_t* w = params_on_device ? mlp_host.data() : params;
	to_operand_layout(w, 64, 32, image.data() + tc::W1_OFF); w += 64 * 32;
	to_operand_layout(w, 16, 64, image.data() + tc::W2_OFF); w += 16 * 64;
	to_operand_layout(w, 64, 32, image.data() + tc::W3_OFF); w += 64 * 32;
	to_operand_layout(w, 64, 64, image.data() + tc::W4_OFF); w += 64 * 64;
	to_operand_layout(w, 16, 64, image.data() + tc::W5_OFF); w += 16 * 64;
	CU(cudaMalloc(&c->d_wimage, tc::W_BYTES));
	CU(cudaMemcpy(c->d_wimage, image.data(), tc::W_BYTES, cudaMemcpyHostToDevice));
	CU(cudaMalloc(&c->d_wrow, kMlpParams * 2));
	CU(cudaMemcpy(c->d_wrow, params, kMlpParams * 2, params_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
	CU(cudaMalloc(&c->d_grid, n_grid * 4));
	CU(cudaMemcpy(c->d_grid, params + kMlpParams, n_grid * 4, params_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
	m.grid = c->d_grid;
	for (uint32_t l = 0; l < desc->n_levels; ++l) m.levels[l].base = c->d_grid + (uintptr_t)m.levels[l].base / sizeof(__half2);
	m.w_image = c->d_wimage;
	m.w_rowmajor = c->d_wrow;
	m.n_levels = desc->n_levels;
	for (int g = 0; g < MAX_LEVELS / 2; ++g) {
		uint32_t a = m.levels[2 * g].hashed, b = m.levels[2 * g + 1].hashed;
		m.pair_mode[g] = (a && b) ? 1u : ((!a && !b) ? 0u : 2u);
	}
	c->model = m;
	c->desc = *desc;
	c->has_model = true;
	return NSB_OK;
}

extern "C" NsbStatus nsb_upload_model(NsbContext* c, const NsbModelDesc* desc, const uint16_t* params, uint64_t n_params) {
	return upload_model_impl(c, desc, params, n_params, false);
}
extern "C" NsbStatus nsb_upload_model_dev(NsbContext* c, const NsbModelDesc* desc, const uint16_t* params_dev, uint64_t n_params) {
	return upload_model_impl(c, desc, params_dev, n_params, true);
}

static NsbStatus upload_occupancy_impl(NsbContext* c, const uint8_t* bitfield, uint64_t n_bytes, bool on_device) {
	if (!c || !bitfield) return fail(NSB_ERR_INVALID, "null argument");
	if (n_bytes != NSB_BITFIELD_BYTES) return fail(NSB_ERR_INVALID, "bitfield must be %u bytes (5 x 128^3 / 8)", NSB_BITFIELD_BYTES);
	CU(cudaSetDevice(c->device));
	if (!c->d_bitfield) CU(cudaMalloc(&c->d_bitfield, NSB_BITFIELD_BYTES));
	CU(cudaDeviceSynchronize());
	CU(cudaMemcpy(c->d_bitfield, bitfield, NSB_BITFIELD_BYTES, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
	c->has_occ = true;
	return NSB_OK;
}
extern "C" NsbStatus nsb_upload_occupancy(NsbContext* c, const uint8_t* bitfield, uint64_t n_bytes) { return upload_occupancy_impl(c, bitfield, n_bytes, false); }
extern "C" NsbStatus nsb_upload_occupancy_dev(NsbContext* c, const uint8_t* bitfield_dev, uint64_t n_bytes) { return upload_occupancy_impl(c, bitfield_dev, n_bytes, true); }

static NsbStatus ensure_grid_buffers(NsbContext* c) {
	const uint32_t n_blocks = (GRIDVOL + 256 * 64 - 1) / (256 * 64);
	if (!c->d_density_grid) {
		CU(cudaMalloc(&c->d_density_grid, (size_t)NSB_GRID_CELLS * 4));
		CU(cudaMemset(c->d_density_grid, 0, (size_t)NSB_GRID_CELLS * 4));
	}
	if (!c->d_mean_partial) CU(cudaMalloc(&c->d_mean_partial, n_blocks * sizeof(double)));
	if (!c->d_thresh) CU(cudaMalloc(&c->d_thresh, 2 * sizeof(float)));
	if (!c->d_bitfield) CU(cudaMalloc(&c->d_bitfield, NSB_BITFIELD_BYTES));
	return NSB_OK;
}
// Testbed::update_density_grid_mean_and_bitfield (testbed_nerf.cu:3642-3658), device-resident and stream-ordered
static NsbStatus rebuild_bitfield(NsbContext* c, cudaStream_t st) {
	const uint32_t n0 = GRIDVOL, n_blocks = (n0 + 256 * 64 - 1) / (256 * 64);
	k_grid_mean_partial<<<n_blocks, 256, 0, st>>>(c->d_density_grid, n0, c->d_mean_partial);  // reduce_sum(fmaxf(val,0)/n) over cascade 0 (:3650)
	k_grid_mean_final<<<1, 1, 0, st>>>(c->d_mean_partial, n_blocks, n0, c->d_thresh);
	k_grid_to_bitfield_dev<<<(NSB_BITFIELD_BYTES + 255) / 256, 256, 0, st>>>(NSB_BITFIELD_BYTES, c->d_density_grid, c->d_bitfield, c->d_thresh);
	for (uint32_t level = 1; level < NSB_NERF_CASCADES; ++level) {
		const uint32_t n = GRIDVOL / 64;
		k_bitfield_max_pool<<<(n + 255) / 256, 256, 0, st>>>(n, c->d_bitfield + (size_t)(level - 1) * (GRIDVOL / 8), c->d_bitfield + (size_t)level * (GRIDVOL / 8));
	}
	CU(cudaGetLastError());
	c->has_occ = true;
	return NSB_OK;
}

extern "C" NsbStatus nsb_upload_density_grid(NsbContext* c, const float* grid, uint64_t n_floats, uint8_t* bitfield_out) {
	if (!c || !grid) return fail(NSB_ERR_INVALID, "null argument");
	if (n_floats != (uint64_t)NSB_GRID_CELLS) return fail(NSB_ERR_INVALID, "density grid must be 5 x 128^3 floats");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	NsbStatus st = ensure_grid_buffers(c);
	if (st != NSB_OK) return st;
	CU(cudaMemcpy(c->d_density_grid, grid, n_floats * 4, cudaMemcpyHostToDevice));
	st = rebuild_bitfield(c, nullptr);
	if (st != NSB_OK) return st;
	CU(cudaDeviceSynchronize());
	if (bitfield_out) CU(cudaMemcpy(bitfield_out, c->d_bitfield, NSB_BITFIELD_BYTES, cudaMemcpyDeviceToHost));
	return NSB_OK;
}

extern "C" NsbStatus nsb_update_density_grid(NsbContext* c, const NsbGridUpdate* u, void* stream_) {
	if (!c || !u) return fail(NSB_ERR_INVALID, "null argument");
	if (!c->has_model) return fail(NSB_ERR_STATE, "nsb_upload_model has not been called");
	if (u->n_cascades < 1 || u->n_cascades > NSB_NERF_CASCADES) return fail(NSB_ERR_INVALID, "n_cascades must be in [1, 5]");
	const uint64_t n_total64 = (uint64_t)u->n_uniform_samples + u->n_nonuniform_samples;
	if (n_total64 > 0xffffff00ull) return fail(NSB_ERR_INVALID, "too many grid samples");
	CU(cudaSetDevice(c->device));
	cudaStream_t st = (cudaStream_t)stream_;
	NsbStatus rc = ensure_grid_buffers(c);
	if (rc != NSB_OK) return rc;
	if (!c->d_density_tmp) CU(cudaMalloc(&c->d_density_tmp, (size_t)NSB_GRID_CELLS * 4));
	if (u->reset_grid) CU(cudaMemsetAsync(c->d_density_grid, 0, (size_t)NSB_GRID_CELLS * 4, st));        // :3558-3560
	CU(cudaMemsetAsync(c->d_density_tmp, 0, (size_t)NSB_GRID_CELLS * 4, st));                            // :3563
	const uint32_t n_total = (uint32_t)n_total64;
	if (n_total > 0) {
		GridUpdateArgs a{};
		a.n_uniform = u->n_uniform_samples; a.n_total = n_total; a.n_cascades = (uint32_t)u->n_cascades; a.step = u->ema_step;
		a.rng_state = u->rng_state; a.rng_inc = u->rng_inc;
		for (int k = 0; k < 3; ++k) { a.amin[k] = u->train_aabb_min[k]; a.amax[k] = u->train_aabb_max[k]; }
		a.density_activation = u->density_activation; a.apply_ops = u->apply_operators ? 1 : 0;
		uint32_t grid = (n_total + 127u) / 128u;
		uint32_t cap = (uint32_t)(c->sm_count * c->grid_update_ctas_per_sm);
		if (grid > cap) grid = cap;
		(c->acc16 ? k_density_grid_update<true> : k_density_grid_update<false>)<<<grid, 128, sizeof(tc::TileSmem), st>>>(c->model, c->d_ops, c->n_ops, a, c->d_density_grid, c->d_density_tmp);
	}
	k_ema_grid<<<(NSB_GRID_CELLS + 255) / 256, 256, 0, st>>>((uint32_t)NSB_GRID_CELLS, u->decay, c->d_density_grid, c->d_density_tmp);  // :3634
	CU(cudaGetLastError());
	return rebuild_bitfield(c, st);                                                                      // :3639
}

extern "C" NsbStatus nsb_download_density_grid(NsbContext* c, float* grid_out, uint8_t* bitfield_out) {
	if (!c) return fail(NSB_ERR_INVALID, "null context");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	if (grid_out) {
		if (!c->d_density_grid) return fail(NSB_ERR_STATE, "no density grid: call nsb_upload_density_grid or nsb_update_density_grid first");
		CU(cudaMemcpy(grid_out, c->d_density_grid, (size_t)NSB_GRID_CELLS * 4, cudaMemcpyDeviceToHost));
	}
	if (bitfield_out) {
		if (!c->d_bitfield || !c->has_occ) return fail(NSB_ERR_STATE, "no occupancy bitfield");
		CU(cudaMemcpy(bitfield_out, c->d_bitfield, NSB_BITFIELD_BYTES, cudaMemcpyDeviceToHost));
	}
	return NSB_OK;
}

template <typename T>
static NsbStatus dev_copy(NsbContext* c, const T* host, size_t count, const T** out) {
	*out = nullptr;
	if (!host || count == 0) return NSB_OK;
	void* p = nullptr;
	CU(cudaMalloc(&p, count * sizeof(T)));
	c->op_allocs.push_back(p);
	CU(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
	*out = reinterpret_cast<const T*>(p);
	return NSB_OK;
}
static void cp3(float* d, const float* s) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
static void cpbox(DevAffineBox& d, const NsbAffineBox& s) { cp3(d.mn, s.min); cp3(d.u, s.u); cp3(d.v, s.v); cp3(d.w, s.w); cp3(d.c, s.center); }

extern "C" NsbStatus nsb_set_edit_ops(NsbContext* c, const NsbEditOp* ops, int32_t n) {
	if (!c || n < 0 || (n > 0 && !ops)) return fail(NSB_ERR_INVALID, "bad arguments");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	free_ops(c);
	if (n == 0) return NSB_OK;
	std::vector<DevOp> dev(n);
	int any_poisson = 0;
	for (int i = 0; i < n; ++i) {
		const NsbEditOp& s = ops[i];
		DevOp& d = dev[i];
		memset(&d, 0, sizeof(d));
		d.type = s.type;
		if (s.type == NSB_OP_CAGE) {
			d.copy = s.copy;
			d.apply_poisson = s.apply_poisson;
			d.amp = s.residual_amplitude;
			cp3(d.amin, s.scene_aabb_min); cp3(d.amax, s.scene_aabb_max);
			cp3(d.bmin, s.bbox_min); cp3(d.bmax, s.bbox_max);
			cp3(d.wbmin, s.warped_bbox_min); cp3(d.wbmax, s.warped_bbox_max);
			cp3(d.owbmin, s.original_warped_bbox_min); cp3(d.owbmax, s.original_warped_bbox_max);
			d.n_tets = s.n_tets;
			if (s.n_tets) {
				if (!s.tet_lut_offsets || !s.tets || !s.vertices || !s.original_vertices || !s.original_bitfield || (s.n_lut_idx && !s.tet_lut_idx)) {
					free_ops(c);
					return fail(NSB_ERR_INVALID, "cage operator %d: missing arrays", i);
				}
				NsbStatus st;
				if ((st = dev_copy(c, s.tet_lut_offsets, (size_t)NSB_GRID_CELLS + 1, &d.lut_off)) != NSB_OK) return st;
				if ((st = dev_copy(c, s.tet_lut_idx, (size_t)s.n_lut_idx, &d.lut_idx)) != NSB_OK) return st;
				if ((st = dev_copy(c, s.tets, (size_t)4 * s.n_tets, &d.tets)) != NSB_OK) return st;
				if ((st = dev_copy(c, s.vertices, (size_t)3 * s.n_vertices, &d.verts)) != NSB_OK) return st;
				if ((st = dev_copy(c, s.original_vertices, (size_t)3 * s.n_vertices, &d.orig_verts)) != NSB_OK) return st;
				if ((st = dev_copy(c, s.local_rotations, (size_t)9 * s.n_tets, &d.rots)) != NSB_OK) return st;
				if ((st = dev_copy(c, s.original_bitfield, (size_t)NSB_BITFIELD_BYTES, &d.obits)) != NSB_OK) return st;
				if (s.boundary_shs && s.boundary_outside_density && s.boundary_residual_density) {
					if ((st = dev_copy(c, s.boundary_shs, (size_t)27 * s.n_vertices, &d.shs)) != NSB_OK) return st;
					if ((st = dev_copy(c, s.boundary_outside_density, (size_t)s.n_vertices, &d.od)) != NSB_OK) return st;
					if ((st = dev_copy(c, s.boundary_residual_density, (size_t)s.n_vertices, &d.rd)) != NSB_OK) return st;
					d.has_poisson_data = 1;
				}
				if (d.apply_poisson && d.has_poisson_data) any_poisson = 1;
			}
		} else if (s.type == NSB_OP_AFFINE) {
			cpbox(d.sel, s.selection_box);
			cpbox(d.dst, s.destination_box);
			cp3(d.translation, s.translation);
			cp3(d.scale, s.scale);
			memcpy(d.rot, s.rotation, sizeof(d.rot));
			d.hide_original = s.hide_original;
			d.correct_dir = s.correct_dir;
		} else {
			free_ops(c);
			return fail(NSB_ERR_INVALID, "operator %d: unknown type %d", i, s.type);
		}
	}
	CU(cudaMalloc(&c->d_ops, sizeof(DevOp) * n));
	CU(cudaMemcpy(c->d_ops, dev.data(), sizeof(DevOp) * n, cudaMemcpyHostToDevice));
	c->n_ops = n;
	c->any_poisson = any_poisson;
	c->h_ops = dev;
	c->rb.assign((size_t)n, NsbContext::CageRebuild{});
	for (int i = 0; i < n; ++i) { c->rb[i].n_vertices = ops[i].type == NSB_OP_CAGE ? ops[i].n_vertices : 0; c->rb[i].n_idx = c->rb[i].idx_cap = ops[i].type == NSB_OP_CAGE ? ops[i].n_lut_idx : 0; }
	return NSB_OK;
}

// ---- per-edit rebuild on the device ---------------------------------------------------------------
static NsbStatus cage_slot(NsbContext* c, int32_t op_index, const char* who) {
	if (!c) return fail(NSB_ERR_INVALID, "%s: null context", who);
	if (op_index < 0 || op_index >= c->n_ops) return fail(NSB_ERR_INVALID, "%s: operator index %d out of range (%d uploaded)", who, op_index, c->n_ops);
	if (c->h_ops[op_index].type != NSB_OP_CAGE || c->h_ops[op_index].n_tets == 0) return fail(NSB_ERR_INVALID, "%s: operator %d is not a cage with a tet mesh", who, op_index);
	return NSB_OK;
}
extern "C" NsbStatus nsb_cage_attach_mvc(NsbContext* c, int32_t op_index, const float* weights, uint32_t n_cage_vertices) {
	NsbStatus st = cage_slot(c, op_index, "nsb_cage_attach_mvc");
	if (st != NSB_OK) return st;
	if (!weights || n_cage_vertices == 0 || n_cage_vertices > 4096) return fail(NSB_ERR_INVALID, "nsb_cage_attach_mvc: bad weights / cage size");
	CU(cudaSetDevice(c->device));
	auto& r = c->rb[op_index];
	cudaFree(r.d_mvc); cudaFree(r.d_cage);
	r.d_mvc = nullptr; r.d_cage = nullptr;
	r.n_cv = n_cage_vertices;
	CU(cudaMalloc(&r.d_mvc, (size_t)r.n_vertices * n_cage_vertices * sizeof(float)));
	CU(cudaMalloc(&r.d_cage, (size_t)3 * n_cage_vertices * sizeof(float)));
	CU(cudaMemcpy(r.d_mvc, weights, (size_t)r.n_vertices * n_cage_vertices * sizeof(float), cudaMemcpyHostToDevice));
	return NSB_OK;
}
extern "C" NsbStatus nsb_cage_deform(NsbContext* c, int32_t op_index, const float* cage_vertices, uint32_t n_cage_vertices, void* stream_) {
	using namespace nsb::rebuild;
	NsbStatus st = cage_slot(c, op_index, "nsb_cage_deform");
	if (st != NSB_OK) return st;
	auto& r = c->rb[op_index];
	if (!r.d_mvc) return fail(NSB_ERR_STATE, "nsb_cage_deform: nsb_cage_attach_mvc has not been called for operator %d", op_index);
	if (!cage_vertices || n_cage_vertices != r.n_cv) return fail(NSB_ERR_INVALID, "nsb_cage_deform: expected %u cage vertices", r.n_cv);
	CU(cudaSetDevice(c->device));
	cudaStream_t s = (cudaStream_t)stream_;
	DevOp& d = c->h_ops[op_index];
	const uint32_t n_cells = (uint32_t)NSB_GRID_CELLS, n_scan_blocks = (n_cells + SCAN_TILE - 1) / SCAN_TILE;
	if (!r.d_counts) {
		CU(cudaMalloc(&r.d_counts, (size_t)n_cells * 4));
		CU(cudaMalloc(&r.d_block_sums, (size_t)n_scan_blocks * 4));
		CU(cudaMalloc(&r.d_total, 4));
		CU(cudaMalloc(&r.d_n_marks, 4));
		CU(cudaMalloc(&r.d_boxes, 18 * sizeof(float)));
		float aabb[6] = {d.amin[0], d.amin[1], d.amin[2], d.amax[0], d.amax[1], d.amax[2]};
		CU(cudaMemcpy(r.d_boxes + 12, aabb, sizeof(aabb), cudaMemcpyHostToDevice));
		r.marks_cap = 1u << 20;
		CU(cudaMalloc(&r.d_marks, (size_t)r.marks_cap * sizeof(Mark)));
	}
	float* d_verts = const_cast<float*>(d.verts);
	const bool verbose = getenv("NSB_VERBOSE") != nullptr;
	auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
	const double t_start = now();
	double t_marked = 0, t_launched = 0;
	// 1. Cage::interpolate_with_mvc
	CU(cudaMemcpyAsync(r.d_cage, cage_vertices, (size_t)3 * r.n_cv * sizeof(float), cudaMemcpyHostToDevice, s));
	k_mvc_apply<<<(r.n_vertices + 127) / 128, 128, 3 * r.n_cv * sizeof(float), s>>>(r.d_mvc, r.n_vertices, r.n_cv, r.d_cage, d_verts);
	// 2. TetMesh::post_update_vertices, update_local_rotations
	k_cage_bbox<<<1, 256, 0, s>>>(d_verts, r.n_vertices, r.d_boxes + 12, r.d_boxes);
	if (d.rots) k_local_rotations<<<(d.n_tets + 63) / 64, 64, 0, s>>>(d_verts, d.orig_verts, d.tets, d.n_tets, const_cast<float*>(d.rots));
	// 3. TetMesh::build_tet_grid: mark, scan, scatter, order
	unsigned int n_marks = 0;
	for (int attempt = 0; attempt < 2; ++attempt) {
		CU(cudaMemsetAsync(r.d_counts, 0, (size_t)n_cells * 4, s));
		CU(cudaMemsetAsync(r.d_n_marks, 0, 4, s));
		k_tet_mark<<<d.n_tets * CASCADES, 128, 0, s>>>(d_verts, d.tets, d.n_tets, r.d_counts, r.d_marks, r.marks_cap, r.d_n_marks);
		CU(cudaMemcpyAsync(&n_marks, r.d_n_marks, 4, cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));  // one sync per edit: the list length decides the allocation
		t_marked = now();
		if (n_marks <= r.marks_cap) break;
		if (attempt == 1) return fail(NSB_ERR_CUDA, "nsb_cage_deform: mark buffer overflow after regrow");
		cudaFree(r.d_marks);
		r.d_marks = nullptr;
		r.marks_cap = n_marks + n_marks / 4;
		CU(cudaMalloc(&r.d_marks, (size_t)r.marks_cap * sizeof(Mark)));
	}
	uint32_t* d_off = const_cast<uint32_t*>(d.lut_off);
	k_scan_block_sums<<<n_scan_blocks, SCAN_THREADS, 0, s>>>(r.d_counts, n_cells, r.d_block_sums);
	k_scan_sums<<<1, SCAN_THREADS, 0, s>>>(r.d_block_sums, n_scan_blocks, r.d_total);
	k_scan_apply<<<n_scan_blocks, SCAN_THREADS, 0, s>>>(r.d_counts, n_cells, r.d_block_sums, r.d_total, d_off);
	uint32_t* d_idx = const_cast<uint32_t*>(d.lut_idx);
	if ((uint64_t)n_marks > r.idx_cap || !d_idx) {  // the list outgrew the array it lives in (uploaded or ours): take a larger one
		cudaFree(r.d_idx);
		r.d_idx = nullptr;
		r.idx_cap = (uint64_t)n_marks + n_marks / 4 + 1;
		CU(cudaMalloc(&r.d_idx, r.idx_cap * 4));
		d_idx = r.d_idx;
		d.lut_idx = r.d_idx;
	}
	if (n_marks) {
		k_tet_fill<<<(n_marks + 255) / 256, 256, 0, s>>>(r.d_marks, n_marks, d_off, r.d_counts, d_idx);
		// the counts are all zero again and the marks are consumed: reuse them as the long-cell queue and the sort scratch
		CU(cudaMemsetAsync(r.d_n_marks, 0, 4, s));
		k_tet_sort<<<(n_cells + 255) / 256, 256, 0, s>>>(d_off, n_cells, d_idx, r.d_counts, r.d_n_marks);
		k_tet_sort_long<<<c->sm_count * 4, 256, 0, s>>>(d_off, d_idx, r.d_counts, r.d_n_marks, reinterpret_cast<uint32_t*>(r.d_marks));
	}
	r.n_idx = n_marks;
	// 4. the operator's boxes and (possibly) its list pointer
	float boxes[12];
	t_launched = now();
	CU(cudaMemcpyAsync(boxes, r.d_boxes, sizeof(boxes), cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	if (verbose)
		fprintf(stderr, "[nsb] cage_deform op %d: %u marks; mark phase %.3f ms, launches %.3f ms, drain %.3f ms\n", op_index, n_marks, t_marked - t_start,
		        t_launched - t_marked, now() - t_launched);
	cp3(d.bmin, boxes); cp3(d.bmax, boxes + 3); cp3(d.wbmin, boxes + 6); cp3(d.wbmax, boxes + 9);
	CU(cudaMemcpyAsync(c->d_ops + op_index, &d, sizeof(DevOp), cudaMemcpyHostToDevice, s));
	CU(cudaGetLastError());
	return NSB_OK;
}
// ---- membrane boundary values ---------------------------------------------------------------------
// Host part of GrowingSelection::compute_poisson_boundary (growing_selection.cu:2235-2260): stratified directions on the sphere,
// jittered by a seeded pcg32 instead of the reference's unseeded std::rand(). libm calls as the reference makes them.
static void boundary_coords(const float* points, uint32_t n_points, const NsbBoundarySampling* p, std::vector<float>& coords) {
	const uint32_t w = p->sampling_width, n_sh = w * w;
	coords.assign((size_t)n_points * n_sh * 7, 0.0f);
	Pcg32 rng;
	rng.state = 0u; rng.inc = (0xda3e39cb94b95bdbULL << 1u) | 1u;  // pcg32(seed): tcnn default stream
	rng.next_uint(); rng.state += p->seed; rng.next_uint();
	auto next_float = [&rng] { uint32_t u = (rng.next_uint() >> 9) | 0x3f800000u; float f; memcpy(&f, &u, 4); return f - 1.0f; };
	for (uint32_t k = 0; k < n_points; ++k)
		for (uint32_t i = 0; i < w; ++i)
			for (uint32_t j = 0; j < w; ++j) {
				const float u = ((float)i + next_float()) / (float)p->hemisphere_width;
				const float v = ((float)j + next_float()) / (float)p->hemisphere_width;
				const float theta = (float)(2.f * M_PI * v);
				const float phi = acosf(2.f * u - 1.f);
				const float x = cosf(theta) * sinf(phi), y = sinf(theta) * sinf(phi), z = cosf(phi);
				float* c = coords.data() + ((size_t)n_sh * k + (size_t)i * w + j) * 7;
				for (int a = 0; a < 3; ++a) c[a] = (points[3 * k + a] - p->train_aabb_min[a]) / (p->train_aabb_max[a] - p->train_aabb_min[a]);  // warp_position
				c[4] = (x + 1.0f) * 0.5f; c[5] = (y + 1.0f) * 0.5f; c[6] = (z + 1.0f) * 0.5f;                                                   // warp_direction
			}
}
extern "C" NsbStatus nsb_poisson_boundary(NsbContext* c, const float* points, uint32_t n_points, const NsbBoundarySampling* p, float* density_out, float* shs_out) {
	if (!c || !points || !p || !density_out || !shs_out) return fail(NSB_ERR_INVALID, "nsb_poisson_boundary: null argument");
	if (!c->has_model) return fail(NSB_ERR_STATE, "nsb_upload_model has not been called");
	if (p->sampling_width == 0 || p->sampling_width > 64 || p->hemisphere_width == 0) return fail(NSB_ERR_INVALID, "nsb_poisson_boundary: bad sampling widths");
	if (p->is_inside && !c->has_occ) return fail(NSB_ERR_STATE, "nsb_poisson_boundary: is_inside needs an occupancy grid");
	if (n_points == 0) return NSB_OK;
	CU(cudaSetDevice(c->device));
	const uint32_t n_sh = p->sampling_width * p->sampling_width;
	const uint64_t n64 = (uint64_t)n_points * n_sh;
	if (n64 > 0x7fffff00ull) return fail(NSB_ERR_INVALID, "nsb_poisson_boundary: too many samples");
	const uint32_t n = (uint32_t)n64, n_padded = next_multiple_u32(n, 128u);  // tcnn::batch_size_granularity (:2231)
	std::vector<float> coords;
	boundary_coords(points, n_points, p, coords);
	float *d_coords = nullptr, *d_aabb = nullptr, *d_dens = nullptr, *d_shs = nullptr;
	__half* d_net = nullptr;
	auto cleanup = [&] { cudaFree(d_coords); cudaFree(d_aabb); cudaFree(d_dens); cudaFree(d_shs); cudaFree(d_net); };
#define CUB_(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail(NSB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); } } while (0)
	CUB_(cudaMalloc(&d_coords, (size_t)n * 7 * 4));
	CUB_(cudaMalloc(&d_aabb, 6 * 4));
	CUB_(cudaMalloc(&d_dens, (size_t)n_points * 4));
	CUB_(cudaMalloc(&d_shs, (size_t)n_points * 27 * 4));
	CUB_(cudaMalloc(&d_net, (size_t)16 * n_padded * 2));
	CUB_(cudaMemcpy(d_coords, coords.data(), (size_t)n * 7 * 4, cudaMemcpyHostToDevice));
	float aabb[6] = {p->train_aabb_min[0], p->train_aabb_min[1], p->train_aabb_min[2], p->train_aabb_max[0], p->train_aabb_max[1], p->train_aabb_max[2]};
	CUB_(cudaMemcpy(d_aabb, aabb, sizeof(aabb), cudaMemcpyHostToDevice));
	uint32_t grid = n_padded / 128, cap = (uint32_t)(c->sm_count * c->inference_ctas_per_sm);
	if (grid > cap) grid = cap;
	(c->acc16 ? k_inference<false, true> : k_inference<false, false>)<<<grid, 128, sizeof(tc::TileSmem)>>>(c->model, d_coords, n, d_net, n_padded);  // inference_mixed_precision (:2286)
	const float scale = (float)(4 * M_PI / (double)n_sh);                                                // :2341
	nsb::rebuild::k_boundary_fit<<<n_points, 32>>>(d_net, n_padded, d_coords, n_sh, p->rgb_activation, p->density_activation,
	                                               p->is_inside ? c->d_bitfield : nullptr, d_aabb, scale, d_dens, d_shs);
	CUB_(cudaGetLastError());
	CUB_(cudaMemcpy(density_out, d_dens, (size_t)n_points * 4, cudaMemcpyDeviceToHost));
	CUB_(cudaMemcpy(shs_out, d_shs, (size_t)n_points * 27 * 4, cudaMemcpyDeviceToHost));
#undef CUB_
	cleanup();
	return NSB_OK;
}
extern "C" NsbStatus nsb_cage_set_membrane(NsbContext* c, int32_t op_index, const float* gamma, uint32_t n_cv, const float* inside_density,
                                           const float* outside_density, const float* inside_shs, const float* outside_shs, float residual_amplitude,
                                           int32_t apply_poisson, float* b_shs_out, float* b_od_out, float* b_rd_out) {
	NsbStatus st = cage_slot(c, op_index, "nsb_cage_set_membrane");
	if (st != NSB_OK) return st;
	if (!gamma || !inside_density || !outside_density || !inside_shs || !outside_shs || n_cv == 0) return fail(NSB_ERR_INVALID, "nsb_cage_set_membrane: null argument");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	DevOp& d = c->h_ops[op_index];
	const uint32_t nv = c->rb[op_index].n_vertices;
	// per-cage-vertex weights on the host, with libm like the reference (:2374-2381)
	std::vector<float> alpha_out(n_cv), w_in(n_cv);
	const float min_step = 1.73205080757f / 1024.0f;  // MIN_CONE_STEPSIZE
	for (uint32_t j = 0; j < n_cv; ++j) {
		const float a_out = 1 - expf(-outside_density[j] * min_step), a_in = 1 - expf(-inside_density[j] * min_step);
		const float q = a_in / a_out;
		alpha_out[j] = a_out;
		w_in[j] = (1.f < q) ? 1.f : q;  // std::min(alpha_in / alpha_out, 1.f)
	}
	float *d_gamma = nullptr, *d_small = nullptr;
	const size_t small = (size_t)n_cv * (4 + 54);
	std::vector<float> pack(small);
	memcpy(pack.data(), alpha_out.data(), n_cv * 4);
	memcpy(pack.data() + n_cv, w_in.data(), n_cv * 4);
	memcpy(pack.data() + 2 * (size_t)n_cv, inside_density, n_cv * 4);
	memcpy(pack.data() + 3 * (size_t)n_cv, outside_density, n_cv * 4);
	memcpy(pack.data() + 4 * (size_t)n_cv, inside_shs, (size_t)n_cv * 27 * 4);
	memcpy(pack.data() + 31 * (size_t)n_cv, outside_shs, (size_t)n_cv * 27 * 4);
	auto& r = c->rb[op_index];
	if (r.mem_n_cv != n_cv) {  // scratch of the blend, kept with the operator (a host re-blends after every boundary re-sampling)
		cudaFree(r.d_gamma); cudaFree(r.d_mem_small);
		r.d_gamma = r.d_mem_small = nullptr;
		CU(cudaMalloc(&r.d_gamma, (size_t)nv * n_cv * 4));
		CU(cudaMalloc(&r.d_mem_small, small * 4));
		r.mem_n_cv = n_cv;
	}
	d_gamma = r.d_gamma;
	d_small = r.d_mem_small;
	CU(cudaMemcpy(d_gamma, gamma, (size_t)nv * n_cv * 4, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d_small, pack.data(), small * 4, cudaMemcpyHostToDevice));
	if (!d.shs) { void* q; CU(cudaMalloc(&q, (size_t)nv * 27 * 4)); c->op_allocs.push_back(q); d.shs = (const float*)q; }
	if (!d.od) { void* q; CU(cudaMalloc(&q, (size_t)nv * 4)); c->op_allocs.push_back(q); d.od = (const float*)q; }
	if (!d.rd) { void* q; CU(cudaMalloc(&q, (size_t)nv * 4)); c->op_allocs.push_back(q); d.rd = (const float*)q; }
	dim3 block(32, 4);
	nsb::rebuild::k_membrane_blend<<<(nv + 3) / 4, block>>>(d_gamma, nv, n_cv, d_small, d_small + n_cv, d_small + 2 * (size_t)n_cv, d_small + 3 * (size_t)n_cv,
	                                                        d_small + 4 * (size_t)n_cv, d_small + 31 * (size_t)n_cv, const_cast<float*>(d.shs),
	                                                        const_cast<float*>(d.od), const_cast<float*>(d.rd));
	CU(cudaGetLastError());
	d.has_poisson_data = 1;
	d.apply_poisson = apply_poisson ? 1 : 0;
	d.amp = residual_amplitude;
	CU(cudaMemcpy(c->d_ops + op_index, &d, sizeof(DevOp), cudaMemcpyHostToDevice));
	int any = 0;
	for (const DevOp& o : c->h_ops) if (o.type == NSB_OP_CAGE && o.apply_poisson && o.has_poisson_data && o.n_tets) any = 1;
	c->any_poisson = any;
	if (b_shs_out) CU(cudaMemcpy(b_shs_out, d.shs, (size_t)nv * 27 * 4, cudaMemcpyDeviceToHost));
	if (b_od_out) CU(cudaMemcpy(b_od_out, d.od, (size_t)nv * 4, cudaMemcpyDeviceToHost));
	if (b_rd_out) CU(cudaMemcpy(b_rd_out, d.rd, (size_t)nv * 4, cudaMemcpyDeviceToHost));
	return NSB_OK;
}
extern "C" NsbStatus nsb_cage_download(NsbContext* c, int32_t op_index, float* vertices, float* rotations, uint32_t* lut_offsets, uint32_t* lut_idx,
                                       uint64_t idx_capacity, uint64_t* n_idx, float* boxes) {
	NsbStatus st = cage_slot(c, op_index, "nsb_cage_download");
	if (st != NSB_OK) return st;
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	const DevOp& d = c->h_ops[op_index];
	const auto& r = c->rb[op_index];
	if (vertices) CU(cudaMemcpy(vertices, d.verts, (size_t)3 * r.n_vertices * 4, cudaMemcpyDeviceToHost));
	if (rotations) {
		if (!d.rots) return fail(NSB_ERR_STATE, "nsb_cage_download: operator %d has no local rotations", op_index);
		CU(cudaMemcpy(rotations, d.rots, (size_t)9 * d.n_tets * 4, cudaMemcpyDeviceToHost));
	}
	if (lut_offsets) CU(cudaMemcpy(lut_offsets, d.lut_off, ((size_t)NSB_GRID_CELLS + 1) * 4, cudaMemcpyDeviceToHost));
	if (n_idx) *n_idx = r.n_idx;
	if (lut_idx) {
		if (idx_capacity < r.n_idx) return fail(NSB_ERR_INVALID, "nsb_cage_download: idx buffer too small (%llu needed)", (unsigned long long)r.n_idx);
		if (r.n_idx) CU(cudaMemcpy(lut_idx, d.lut_idx, (size_t)r.n_idx * 4, cudaMemcpyDeviceToHost));
	}
	if (boxes) { cp3(boxes, d.bmin); cp3(boxes + 3, d.bmax); cp3(boxes + 6, d.wbmin); cp3(boxes + 9, d.wbmax); }
	return NSB_OK;
}

// ---- frame conversion ---------------------------------------------------------------------------------
static const uint32_t kSobolDir1[32] = {
	0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000, 0xcc000000, 0xaa000000, 0xff000000, 0x80800000, 0xc0c00000, 0xa0a00000,
	0xf0f00000, 0x88880000, 0xcccc0000, 0xaaaa0000, 0xffff0000, 0x80008000, 0xc000c000, 0xa000a000, 0xf000f000, 0x88008800, 0xcc00cc00,
	0xaa00aa00, 0xff00ff00, 0x80808080, 0xc0c0c0c0, 0xa0a0a0a0, 0xf0f0f0f0, 0x88888888, 0xcccccccc, 0xaaaaaaaa, 0xffffffff,
};
// ld_random_val_2d (random_val.cuh:278-282)
static void host_ld_random_val_2d(uint32_t index, uint32_t seed, float* out) {
	index = nested_uniform_scramble_base2(index, seed);
	for (uint32_t dim = 0; dim < 2; ++dim) {
		uint32_t X = 0;
		for (uint32_t bit = 0; bit < 32; ++bit) {
			uint32_t mask = (index >> bit) & 1u;
			X ^= mask * (dim == 0 ? (0x80000000u >> bit) : kSobolDir1[bit]);
		}
		out[dim] = (float)nested_uniform_scramble_base2(X, hash_combine(seed, dim)) * 2.3283064365386963e-10f;
	}
}
// ld_random_pixel_offset (random_val.cuh:317-322): the same for every pixel, so evaluated once here
static void host_pixel_offset(uint32_t spp, float* off) {
	float a[2], b[2];
	host_ld_random_val_2d(0, 0xdeadbeefu, a);
	host_ld_random_val_2d(spp, 0xdeadbeefu, b);
	for (int i = 0; i < 2; ++i) {
		volatile float s = 0.5f - a[i];
		volatile float o = s + b[i];
		off[i] = o - floorf(o);
	}
}

static NsbStatus to_dev_frame(const NsbFrame* f, DevFrame* d) {
	if (!f) return fail(NSB_ERR_INVALID, "frame is null");
	if (f->width <= 0 || f->height <= 0 || (uint64_t)f->width * f->height > 0x7fffffffull / 786433ull * 786433ull) return fail(NSB_ERR_INVALID, "bad resolution %dx%d", f->width, f->height);
	if (f->tile_world <= 0 || f->tile_rank < 0 || f->tile_rank >= f->tile_world) return fail(NSB_ERR_INVALID, "bad tile partition %d/%d", f->tile_rank, f->tile_world);
	switch (f->render_mode) {
		case NSB_RENDER_AO: case NSB_RENDER_SHADE: case NSB_RENDER_POSITIONS: case NSB_RENDER_DEPTH:
		case NSB_RENDER_DISTANCE: case NSB_RENDER_STEPSIZE: case NSB_RENDER_COST: case NSB_RENDER_DISTORTION: break;
		default: return fail(NSB_ERR_UNSUPPORTED, "render mode %d is not covered (Normals/EncodingVis need network input gradients; Slice is a side path)", f->render_mode);
	}
	if (f->camera_distortion_mode < 0 || f->camera_distortion_mode > 2) return fail(NSB_ERR_INVALID, "bad camera distortion mode %d", f->camera_distortion_mode);
	if (f->focus_z < 0.0f) return fail(NSB_ERR_UNSUPPORTED, "negative focus_z selects the reference's Slice side path (testbed_nerf.cu:3067-3070), which is not covered");
	if (f->dof != 0.0f && !(f->focus_z > 0.0f)) return fail(NSB_ERR_INVALID, "depth of field needs focus_z > 0");
	if (f->envmap_dev && (f->envmap_resolution[0] <= 0 || f->envmap_resolution[1] <= 0)) return fail(NSB_ERR_INVALID, "bad envmap resolution");
	if (f->distortion_dev && (f->distortion_resolution[0] <= 0 || f->distortion_resolution[1] <= 0)) return fail(NSB_ERR_INVALID, "bad distortion map resolution");
	if (f->rgb_activation < 0 || f->rgb_activation > 3 || f->density_activation < 0 || f->density_activation > 3) return fail(NSB_ERR_INVALID, "bad activation");
	d->W = f->width; d->H = f->height;
	d->fx = f->focal_length[0]; d->fy = f->focal_length[1];
	d->cx = f->screen_center[0]; d->cy = f->screen_center[1];
	memcpy(d->cam0, f->camera0, sizeof(d->cam0));
	memcpy(d->cam1, f->camera1, sizeof(d->cam1));
	memcpy(d->rs, f->rolling_shutter, sizeof(d->rs));
	cp3(d->rmin, f->render_aabb_min); cp3(d->rmax, f->render_aabb_max);
	cp3(d->tmin, f->train_aabb_min); cp3(d->tmax, f->train_aabb_max);
	d->cone = f->cone_angle_constant;
	d->min_T = f->min_transmittance;
	d->depth_scale = f->depth_scale;
	d->rgb_act = f->rgb_activation; d->density_act = f->density_activation; d->mode = f->render_mode;
	d->spp = f->spp_index;
	host_pixel_offset(f->snap_to_pixel_centers ? 0u : f->spp_index, d->pix_off);
	d->apply_ops = f->apply_operators; d->poisson_target = f->poisson_target; d->linear_colors = f->linear_colors;
	d->min_mip = f->min_mip < 0 ? 0 : (f->min_mip > 4 ? 4 : f->min_mip);
	d->show_accel = f->show_accel != 0;
	d->tile_rank = f->tile_rank; d->tile_world = f->tile_world;
	d->tiles_x = (f->width + TILE_W - 1) / TILE_W;
	d->tiles_y = (f->height + TILE_H - 1) / TILE_H;
	d->cam_dist_mode = f->camera_distortion_mode;
	memcpy(d->cam_dist, f->camera_distortion_params, sizeof(d->cam_dist));
	d->dof = f->dof; d->focus_z = f->focus_z;
	d->glow_mode = f->glow_mode; d->glow_y_cutoff = f->glow_y_cutoff;
	d->envmap = f->envmap_dev; d->env_w = f->envmap_resolution[0]; d->env_h = f->envmap_resolution[1];
	d->distortion = f->distortion_dev; d->dist_w = f->distortion_resolution[0]; d->dist_h = f->distortion_resolution[1];
	d->general_camera = (d->cam_dist_mode != 0 || d->distortion != nullptr || d->dof != 0.0f) ? 1 : 0;
	return NSB_OK;
}

extern "C" NsbStatus nsb_render(NsbContext* c, const NsbFrame* frame, float* fb_dev, float* depth_dev, void* stream_) {
	if (!c || !fb_dev || !depth_dev) return fail(NSB_ERR_INVALID, "null argument");
	if (!c->has_model) return fail(NSB_ERR_STATE, "nsb_upload_model has not been called");
	DevFrame f;
	NsbStatus st = to_dev_frame(frame, &f);
	if (st != NSB_OK) return st;
	CU(cudaSetDevice(c->device));
	cudaStream_t stream = (cudaStream_t)stream_;
	uint32_t n_tiles = (uint32_t)(f.tiles_x * f.tiles_y);
	uint32_t my_tiles = (n_tiles + (uint32_t)f.tile_world - 1 - (uint32_t)f.tile_rank) / (uint32_t)f.tile_world;
	size_t n_local = (size_t)my_tiles * TILE_PIXELS;
	if (n_local > c->list_capacity) {
		CU(cudaStreamSynchronize(stream));
		cudaFree(c->d_list);
		c->d_list = nullptr; c->list_capacity = 0;
		CU(cudaMalloc(&c->d_list, n_local * sizeof(RayRec)));
		c->list_capacity = n_local;
	}
	// frames whose ray generation is not the pinhole fast path (ABI 3): k_prepare_rays<true>; with a general camera the rays travel in a side buffer
	const bool general = f.general_camera || f.envmap != nullptr || f.mode == NSB_RENDER_DISTORTION;
	float* ray_od = nullptr;
	if (f.general_camera) {
		if (n_local > c->ray_od_capacity) {
			CU(cudaStreamSynchronize(stream));
			cudaFree(c->d_ray_od);
			c->d_ray_od = nullptr; c->ray_od_capacity = 0;
			CU(cudaMalloc(&c->d_ray_od, n_local * 6 * sizeof(float)));
			c->ray_od_capacity = n_local;
		}
		ray_od = c->d_ray_od;
	}
	CU(cudaEventRecord(c->ev0, stream));
	CU(cudaEventRecord(c->evm, stream));
	CU(cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(uint32_t), stream));
	CU(cudaMemsetAsync(c->d_stats, 0, ST_N * sizeof(unsigned long long), stream));
	if (n_local > 0) {
		if (general)
			k_prepare_rays<true><<<(unsigned)((n_local + 255) / 256), 256, 0, stream>>>(f, c->has_occ ? c->d_bitfield : nullptr, reinterpret_cast<float4*>(fb_dev), depth_dev, c->d_list,
			                                                                         ray_od, c->d_counters, (uint32_t)n_local, c->d_stats);
		else
			k_prepare_rays<false><<<(unsigned)((n_local + 255) / 256), 256, 0, stream>>>(f, c->has_occ ? c->d_bitfield : nullptr, reinterpret_cast<float4*>(fb_dev), depth_dev, c->d_list,
			                                                                          nullptr, c->d_counters, (uint32_t)n_local, c->d_stats);
		CU(cudaGetLastError());
		CU(cudaEventRecord(c->evm, stream));
		const bool ops_on = f.apply_ops && c->n_ops > 0;
		const uint32_t tiles = (uint32_t)render_tiles(ops_on);
		uint32_t grid = (uint32_t)(c->sm_count * c->ctas_per_sm);
		if (grid > (my_tiles + tiles - 1) / tiles) grid = (my_tiles + tiles - 1) / tiles;
		if (!ops_on && c->use_ws && !c->acc16 && !f.show_accel && !f.general_camera && !f.glow_mode && (uint64_t)f.W * (uint64_t)f.H < (1ull << 28)) {  // k_render_ws ships the pixel in 28 bits of a flag word
			uint32_t g2 = (uint32_t)c->sm_count;
			if (g2 > (my_tiles + ws::PT - 1) / ws::PT) g2 = (uint32_t)((my_tiles + ws::PT - 1) / ws::PT);
			k_render_ws<<<g2, ws::THREADS, sizeof(ws::Smem), stream>>>(f, c->model, c->has_occ ? c->d_bitfield : nullptr, reinterpret_cast<float4*>(fb_dev), depth_dev, c->d_list,
			                                                           c->d_counters, c->d_counters + 1, c->d_stats, c->refill_thr, c->dda_budget);
		} else {
		auto kernel = ops_on ? (c->acc16 ? k_render_fused<true, true> : k_render_fused<true, false>) : (c->acc16 ? k_render_fused<false, true> : k_render_fused<false, false>);
		kernel<<<grid, 128 * tiles, ops_on ? sizeof(RenderSmemT<render_tiles(true)>) : sizeof(RenderSmemT<render_tiles(false)>), stream>>>(f, c->model, c->has_occ ? c->d_bitfield : nullptr, c->d_ops, c->n_ops, c->any_poisson,
		                                                          reinterpret_cast<float4*>(fb_dev), depth_dev, c->d_list, ray_od, c->d_counters, c->d_counters + 1, c->d_stats,
		                                                          c->refill_thr, c->dda_budget | (c->helpers << 16));
		}
		CU(cudaGetLastError());
	}
	CU(cudaEventRecord(c->ev1, stream));
	c->timed = true;
	c->launches = 2;
	return NSB_OK;
}

extern "C" NsbStatus nsb_render_host(NsbContext* c, const NsbFrame* frame, float* fb_host, float* depth_host) {
	if (!c || !frame || !fb_host) return fail(NSB_ERR_INVALID, "null argument");
	CU(cudaSetDevice(c->device));
	size_t n = (size_t)frame->width * (size_t)frame->height;
	if (n == 0) return fail(NSB_ERR_INVALID, "empty frame");
	if (n > c->fb_pixels) {
		cudaFree(c->d_fb); cudaFree(c->d_depth);
		c->d_fb = nullptr; c->d_depth = nullptr; c->fb_pixels = 0;
		CU(cudaMalloc(&c->d_fb, n * 16));
		CU(cudaMalloc(&c->d_depth, n * 4));
		c->fb_pixels = n;
	}
	CU(cudaMemsetAsync(c->d_fb, 0, n * 16, c->stream));  // render_buffer.clear_frame (testbed.cu:2635)
	CU(cudaMemsetAsync(c->d_depth, 0, n * 4, c->stream));
	NsbStatus st = nsb_render(c, frame, c->d_fb, c->d_depth, c->stream);
	if (st != NSB_OK) return st;
	CU(cudaMemcpyAsync(fb_host, c->d_fb, n * 16, cudaMemcpyDeviceToHost, c->stream));
	if (depth_host) CU(cudaMemcpyAsync(depth_host, c->d_depth, n * 4, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return NSB_OK;
}

extern "C" NsbStatus nsb_render_host_async(NsbContext* c, const NsbFrame* frame, float* fb_host, float* depth_host, uint64_t* ticket) {
	if (!c || !frame || !fb_host || !ticket) return fail(NSB_ERR_INVALID, "null argument");
	CU(cudaSetDevice(c->device));
	const size_t n = (size_t)frame->width * (size_t)frame->height;
	if (n == 0) return fail(NSB_ERR_INVALID, "empty frame");
	if (n > c->fb2_pixels) {
		CU(cudaStreamSynchronize(c->stream));
		CU(cudaStreamSynchronize(c->copy_stream));
		for (int i = 0; i < 2; ++i) {
			cudaFree(c->d_fb2[i]); cudaFree(c->d_depth2[i]);
			c->d_fb2[i] = nullptr; c->d_depth2[i] = nullptr;
		}
		c->fb2_pixels = 0;
		for (int i = 0; i < 2; ++i) {
			CU(cudaMalloc(&c->d_fb2[i], n * 16));
			CU(cudaMalloc(&c->d_depth2[i], n * 4));
		}
		c->fb2_pixels = n;
	}
	const int slot = (int)(c->host_seq & 1u);
	if (c->host_seq >= 2) CU(cudaStreamWaitEvent(c->stream, c->copy_done[slot], 0));  // the copy that last read this slot (two tickets ago)
	CU(cudaMemsetAsync(c->d_fb2[slot], 0, n * 16, c->stream));  // render_buffer.clear_frame (testbed.cu:2635)
	CU(cudaMemsetAsync(c->d_depth2[slot], 0, n * 4, c->stream));
	NsbStatus st = nsb_render(c, frame, c->d_fb2[slot], c->d_depth2[slot], c->stream);
	if (st != NSB_OK) return st;
	CU(cudaEventRecord(c->render_done[slot], c->stream));
	CU(cudaStreamWaitEvent(c->copy_stream, c->render_done[slot], 0));
	CU(cudaMemcpyAsync(fb_host, c->d_fb2[slot], n * 16, cudaMemcpyDeviceToHost, c->copy_stream));
	if (depth_host) CU(cudaMemcpyAsync(depth_host, c->d_depth2[slot], n * 4, cudaMemcpyDeviceToHost, c->copy_stream));
	CU(cudaEventRecord(c->copy_done[slot], c->copy_stream));
	*ticket = c->host_seq++;
	return NSB_OK;
}

extern "C" NsbStatus nsb_host_frame_wait(NsbContext* c, uint64_t ticket) {
	if (!c) return fail(NSB_ERR_INVALID, "null context");
	if (ticket >= c->host_seq) return fail(NSB_ERR_INVALID, "ticket %llu was never issued", (unsigned long long)ticket);
	CU(cudaSetDevice(c->device));
	CU(cudaEventSynchronize(c->copy_done[ticket & 1u]));  // if the slot was reused since, this is the later copy, which started after this one finished
	return NSB_OK;
}

extern "C" NsbStatus nsb_set_mlp_accumulator(NsbContext* c, int32_t policy) {
	if (!c) return fail(NSB_ERR_INVALID, "null context");
	if (policy != NSB_MLP_ACC_F32 && policy != NSB_MLP_ACC_F16) return fail(NSB_ERR_INVALID, "unknown accumulator policy %d", policy);
	c->acc16 = policy == NSB_MLP_ACC_F16;
	return NSB_OK;
}

extern "C" NsbStatus nsb_get_stats(NsbContext* c, NsbRenderStats* out) {
	if (!c || !out) return fail(NSB_ERR_INVALID, "null argument");
	CU(cudaSetDevice(c->device));
	memset(out, 0, sizeof(*out));
	if (!c->timed) return NSB_OK;
	CU(cudaEventSynchronize(c->ev1));
	unsigned long long h[ST_N];
	CU(cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
	out->n_rays = h[ST_RAYS]; out->n_rays_alive = h[ST_ALIVE]; out->n_hit = h[ST_HIT]; out->n_samples = h[ST_SAMPLES]; out->n_old_samples = h[ST_OLD];
	out->n_kernel_launches = c->launches;
	CU(cudaEventElapsedTime(&out->gpu_ms, c->ev0, c->ev1));
	CU(cudaEventElapsedTime(&out->fused_ms, c->evm, c->ev1));
	return NSB_OK;
}

// Diagnostic counters of the last render: rounds (CTA x MLP rounds), and thread-0 cycles per phase summed over CTAs:
// {rounds, acquire, encode, mlp, composite, total, ctas}
extern "C" NsbStatus nsb_debug_counters(NsbContext* c, uint64_t* out, int32_t n) {
	if (!c || !out || n < 7) return fail(NSB_ERR_INVALID, "need room for 7 counters");
	CU(cudaSetDevice(c->device));
	if (c->timed) CU(cudaEventSynchronize(c->ev1));
	unsigned long long h[ST_N];
	CU(cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
	for (int i = 0; i < 7; ++i) out[i] = h[ST_ROUNDS + i];
	return NSB_OK;
}

extern "C" NsbStatus nsb_tiles_for_rank(int32_t width, int32_t height, int32_t rank, int32_t world, uint32_t* n_tiles) {
	if (!n_tiles || width <= 0 || height <= 0 || world <= 0 || rank < 0 || rank >= world) return fail(NSB_ERR_INVALID, "bad arguments");
	uint32_t total = (uint32_t)(((width + TILE_W - 1) / TILE_W) * ((height + TILE_H - 1) / TILE_H));
	*n_tiles = (total + (uint32_t)world - 1 - (uint32_t)rank) / (uint32_t)world;
	return NSB_OK;
}

static NsbStatus pack_impl(NsbContext* c, const float* fb, const float* depth, int W, int H, int rank, int world, float* pr, float* pd, void* stream, int unpack) {
	if (!c || !fb || !pr) return fail(NSB_ERR_INVALID, "null argument");
	uint32_t mine = 0;
	NsbStatus st = nsb_tiles_for_rank(W, H, rank, world, &mine);
	if (st != NSB_OK) return st;
	CU(cudaSetDevice(c->device));
	int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
	if (mine == 0) return NSB_OK;
	k_pack_tiles<<<mine, TILE_PIXELS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(fb), depth, W, H, tiles_x, tiles_x * tiles_y, rank, world,
	                                                            reinterpret_cast<float4*>(pr), pd, unpack);
	CU(cudaGetLastError());
	return NSB_OK;
}
extern "C" NsbStatus nsb_pack_tiles(NsbContext* c, const float* fb_dev, const float* depth_dev, int32_t W, int32_t H, int32_t rank, int32_t world,
                                    float* dst_rgba, float* dst_depth, void* stream) {
	return pack_impl(c, fb_dev, depth_dev, W, H, rank, world, dst_rgba, dst_depth, stream, 0);
}
extern "C" NsbStatus nsb_unpack_tiles(NsbContext* c, const float* src_rgba, const float* src_depth, int32_t W, int32_t H, int32_t rank, int32_t world,
                                      float* fb_dev, float* depth_dev, void* stream) {
	return pack_impl(c, fb_dev, depth_dev, W, H, rank, world, const_cast<float*>(src_rgba), const_cast<float*>(src_depth), stream, 1);
}

extern "C" NsbStatus nsb_unpack_gathered(NsbContext* c, const float* gathered_rgba, const float* gathered_depth, int32_t W, int32_t H, int32_t skip_rank, int32_t world,
                                         uint32_t rank_stride_floats, float* fb_dev, float* depth_dev, void* stream) {
	if (!c || !gathered_rgba || !fb_dev || W <= 0 || H <= 0 || world <= 0) return fail(NSB_ERR_INVALID, "bad arguments");
	CU(cudaSetDevice(c->device));
	const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
	const uint32_t need = (uint32_t)((tiles_x * tiles_y + world - 1) / world);
	if (rank_stride_floats % 4 != 0 || rank_stride_floats < need * (uint32_t)TILE_PIXELS * 4u) return fail(NSB_ERR_INVALID, "rank stride %u floats is not a multiple of 4 or smaller than rank 0's %u tiles", rank_stride_floats, need);
	k_unpack_gathered<<<dim3(need, (unsigned)world), TILE_PIXELS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(gathered_rgba), gathered_depth, W, H, tiles_x, tiles_x * tiles_y,
	                                                                                         skip_rank, world, rank_stride_floats, reinterpret_cast<float4*>(fb_dev), depth_dev);
	CU(cudaGetLastError());
	return NSB_OK;
}

extern "C" NsbStatus nsb_accumulate(NsbContext* c, const float* frame, float* acc, int32_t W, int32_t H, uint32_t spp, int32_t color_space, void* stream) {
	if (!c || !frame || !acc || W <= 0 || H <= 0) return fail(NSB_ERR_INVALID, "bad arguments");
	if (color_space < 0 || color_space > 2) return fail(NSB_ERR_INVALID, "bad colour space %d", color_space);
	CU(cudaSetDevice(c->device));
	int n = W * H;
	if (spp == 0) CU(cudaMemsetAsync(acc, 0, (size_t)n * 16, (cudaStream_t)stream));  // render_buffer.cu:545-547
	k_accumulate<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, reinterpret_cast<const float4*>(frame), reinterpret_cast<float4*>(acc), (float)spp, color_space);
	CU(cudaGetLastError());
	return NSB_OK;
}
extern "C" NsbStatus nsb_tonemap(NsbContext* c, const float* acc, float* out, int32_t W, int32_t H, const NsbTonemap* p, void* stream) {
	if (!c || !acc || !out || !p || W <= 0 || H <= 0) return fail(NSB_ERR_INVALID, "bad arguments");
	if (p->color_space < 0 || p->color_space > 2 || p->output_color_space < 0 || p->output_color_space > 1 || p->tonemap_curve < 0 || p->tonemap_curve > 3)
		return fail(NSB_ERR_INVALID, "bad tonemap parameters");
	CU(cudaSetDevice(c->device));
	int n = W * H;
	k_tonemap<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, reinterpret_cast<const float4*>(acc), reinterpret_cast<float4*>(out), *p);
	CU(cudaGetLastError());
	return NSB_OK;
}

// ---- operator-level entry points -------------------------------------------------------------------------
template <bool DENSITY_ONLY>
static NsbStatus inference_impl(NsbContext* c, const float* coords, uint32_t n, uint16_t* out, uint32_t n_padded, void* stream) {
	if (!c) return fail(NSB_ERR_INVALID, "null context");
	if (!c->has_model) return fail(NSB_ERR_STATE, "nsb_upload_model has not been called");
	if (n_padded < n || n_padded % 128 != 0) return fail(NSB_ERR_INVALID, "n_padded must be a multiple of 128 (tcnn::batch_size_granularity) and >= n");
	if (n_padded == 0) return NSB_OK;  // empty batch
	if (!coords || !out) return fail(NSB_ERR_INVALID, "null argument");
	CU(cudaSetDevice(c->device));
	uint32_t grid = n_padded / 128;
	uint32_t cap = (uint32_t)(c->sm_count * c->inference_ctas_per_sm);
	if (grid > cap) grid = cap;
	(c->acc16 ? k_inference<DENSITY_ONLY, true> : k_inference<DENSITY_ONLY, false>)<<<grid, 128, sizeof(tc::TileSmem), (cudaStream_t)stream>>>(c->model, coords, n, reinterpret_cast<__half*>(out), n_padded);
	CU(cudaGetLastError());
	return NSB_OK;
}
extern "C" NsbStatus nsb_inference(NsbContext* c, const float* coords, uint32_t n, uint16_t* out, uint32_t n_padded, void* stream) {
	return inference_impl<false>(c, coords, n, out, n_padded, stream);
}
extern "C" NsbStatus nsb_density(NsbContext* c, const float* coords, uint32_t n, uint16_t* out, uint32_t n_padded, void* stream) {
	return inference_impl<true>(c, coords, n, out, n_padded, stream);
}
extern "C" NsbStatus nsb_encode(NsbContext* c, const float* coords, uint32_t n, uint16_t* out, uint32_t n_padded, void* stream) {
	if (!c || !coords || !out) return fail(NSB_ERR_INVALID, "null argument");
	if (!c->has_model) return fail(NSB_ERR_STATE, "nsb_upload_model has not been called");
	if (n_padded < n) return fail(NSB_ERR_INVALID, "n_padded < n");
	if (n == 0) return NSB_OK;
	CU(cudaSetDevice(c->device));
	k_encode<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->model, coords, n, reinterpret_cast<__half*>(out), n_padded);
	CU(cudaGetLastError());
	return NSB_OK;
}
// op_index < 0: every uploaded operator in reverse list order (the loops of testbed_nerf.cu:2868-2904, outputs cleared first like :2863-2866, :2898);
// op_index >= 0: that operator alone = one EditOperator virtual call (outputs are not cleared, as the virtuals do not clear them)
static NsbStatus op_range(NsbContext* c, int32_t op_index, const DevOp** ops, int* n_ops) {
	if (op_index >= c->n_ops) return fail(NSB_ERR_INVALID, "operator %d of %d", op_index, c->n_ops);
	*ops = op_index < 0 ? c->d_ops : c->d_ops + op_index;
	*n_ops = op_index < 0 ? c->n_ops : 1;
	return NSB_OK;
}
extern "C" NsbStatus nsb_map_rays_op(NsbContext* c, int32_t op_index, float* coords, uint8_t* empty_mask, uint32_t n, void* stream) {
	if (!c || !coords || !empty_mask) return fail(NSB_ERR_INVALID, "null argument");
	if (n == 0) return NSB_OK;
	CU(cudaSetDevice(c->device));
	if (op_index < 0) CU(cudaMemsetAsync(empty_mask, 0, n, (cudaStream_t)stream));  // :2898
	if (c->n_ops == 0) return NSB_OK;
	const DevOp* ops; int n_ops;
	NsbStatus st = op_range(c, op_index, &ops, &n_ops);
	if (st != NSB_OK) return st;
	k_map_rays<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ops, n_ops, coords, empty_mask, n);
	CU(cudaGetLastError());
	return NSB_OK;
}
extern "C" NsbStatus nsb_map_rays(NsbContext* c, float* coords, uint8_t* empty_mask, uint32_t n, void* stream) { return nsb_map_rays_op(c, -1, coords, empty_mask, n, stream); }
extern "C" NsbStatus nsb_poisson_residuals_op(NsbContext* c, int32_t op_index, const float* coords, uint32_t n, float* sh, float* od, float* rd, void* stream) {
	if (!c || !coords || !sh || !od || !rd) return fail(NSB_ERR_INVALID, "null argument");
	if (n == 0) return NSB_OK;
	CU(cudaSetDevice(c->device));
	if (op_index < 0) {  // :2863-2866
		CU(cudaMemsetAsync(sh, 0, (size_t)n * 27 * sizeof(float), (cudaStream_t)stream));
		CU(cudaMemsetAsync(od, 0, (size_t)n * sizeof(float), (cudaStream_t)stream));
		CU(cudaMemsetAsync(rd, 0, (size_t)n * sizeof(float), (cudaStream_t)stream));
	}
	if (c->n_ops == 0) return NSB_OK;
	const DevOp* ops; int n_ops;
	NsbStatus st = op_range(c, op_index, &ops, &n_ops);
	if (st != NSB_OK) return st;
	k_poisson<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ops, n_ops, coords, n, sh, od, rd);
	CU(cudaGetLastError());
	return NSB_OK;
}
extern "C" NsbStatus nsb_poisson_residuals(NsbContext* c, const float* coords, uint32_t n, float* sh, float* od, float* rd, void* stream) {
	return nsb_poisson_residuals_op(c, -1, coords, n, sh, od, rd, stream);
}
extern "C" NsbStatus nsb_map_positions(NsbContext* c, int32_t op_index, float* positions, uint32_t stride_floats, uint8_t* empty_mask, uint32_t n, void* stream) {
	if (!c || !positions || stride_floats < 3) return fail(NSB_ERR_INVALID, "bad arguments");
	if (n == 0) return NSB_OK;
	CU(cudaSetDevice(c->device));
	if (op_index < 0 && empty_mask) CU(cudaMemsetAsync(empty_mask, 0, n, (cudaStream_t)stream));  // :3595
	if (c->n_ops == 0) return NSB_OK;
	const DevOp* ops; int n_ops;
	NsbStatus st = op_range(c, op_index, &ops, &n_ops);
	if (st != NSB_OK) return st;
	k_map_positions<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ops, n_ops, positions, stride_floats, empty_mask, n);
	CU(cudaGetLastError());
	return NSB_OK;
}
extern "C" NsbStatus nsb_poisson_residual_density(NsbContext* c, int32_t op_index, const float* positions, uint32_t stride_floats, uint16_t* density_fp16, uint32_t n, void* stream) {
	if (!c || !positions || !density_fp16 || stride_floats < 3) return fail(NSB_ERR_INVALID, "bad arguments");
	if (n == 0 || c->n_ops == 0) return NSB_OK;
	CU(cudaSetDevice(c->device));
	const DevOp* ops; int n_ops;
	NsbStatus st = op_range(c, op_index, &ops, &n_ops);
	if (st != NSB_OK) return st;
	k_poisson_residual_density<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ops, n_ops, positions, stride_floats, reinterpret_cast<__half*>(density_fp16), n);
	CU(cudaGetLastError());
	return NSB_OK;
}
extern "C" NsbStatus nsb_march_trace(NsbContext* c, const NsbFrame* frame, const uint32_t* pixels, uint32_t n_pixels, uint32_t max_samples, float* rec,
                                     uint32_t* idx, uint32_t* count, void* stream) {
	if (!c || !pixels || !rec || !idx || !count) return fail(NSB_ERR_INVALID, "null argument");
	DevFrame f;
	NsbStatus st = to_dev_frame(frame, &f);
	if (st != NSB_OK) return st;
	if (n_pixels == 0) return NSB_OK;
	CU(cudaSetDevice(c->device));
	if (f.general_camera) k_march_trace<true><<<(n_pixels + 63) / 64, 64, 0, (cudaStream_t)stream>>>(f, c->has_occ ? c->d_bitfield : nullptr, pixels, n_pixels, max_samples, rec, idx, count);
	else k_march_trace<false><<<(n_pixels + 63) / 64, 64, 0, (cudaStream_t)stream>>>(f, c->has_occ ? c->d_bitfield : nullptr, pixels, n_pixels, max_samples, rec, idx, count);
	CU(cudaGetLastError());
	return NSB_OK;
}
