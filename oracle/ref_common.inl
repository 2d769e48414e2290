// ref_common.inl — shared by oracle/ref_driver.cpp (CPU build) and oracle/ref_driver_cuda.cu (nvcc build): TEST INFRASTRUCTURE ONLY.
// Declarations of the Testbed / NerfTracer / NerfNetwork / CudaRenderBuffer members that the reference's cut-out functions touch
// (same names as the reference's headers, which cannot be included: they pull in the GUI, training and tiny-cuda-nn), then the
// cut-outs themselves, then the EditOperator subclasses that marshal the C-ABI PODs into the reference's kernels.
// The including file provides nsb_ref_alloc(bytes) -> std::shared_ptr<void> (host or device memory).
// ---------------------------------------------------------------------------------------------------------------------
// declarations the cut-out members need (names = the reference's; only what the render path touches)
// ---------------------------------------------------------------------------------------------------------------------
namespace tcnn {
static size_t g_scratch_floor = 0;
struct GPUMemoryArena {
	struct Allocation { std::vector<std::shared_ptr<void>> blocks; };
};
// tcnn::allocate_workspace_and_distribute<Types...>(stream, &alloc, sizes...) -> tuple<Types*...>
template <typename... Types, size_t... I> std::tuple<Types*...> distribute_impl(GPUMemoryArena::Allocation* alloc, const size_t* sizes, std::index_sequence<I...>) {
	alloc->blocks.clear();
	for (size_t k = 0; k < sizeof...(Types); ++k) alloc->blocks.push_back(nullptr);
	const size_t elem[] = {sizeof(Types)...};
	// g_scratch_floor: the reference's occupancy update launches compute_poisson_residual_density over 5*128^3 elements although its buffers hold
	// n_density_grid_samples (testbed_nerf.cu:3612-3620): on the GPU those are stray reads/writes inside the arena that touch no used value; the CPU
	// build gives them room instead of a segmentation fault.
	for (size_t k = 0; k < sizeof...(Types); ++k) alloc->blocks[k] = nsb_ref_alloc(std::max(sizes[k] * elem[k], g_scratch_floor) + 64);
	return std::tuple<Types*...>{(Types*)alloc->blocks[I].get()...};
}
template <typename... Types, typename... Sizes> std::tuple<Types*...> allocate_workspace_and_distribute(cudaStream_t, GPUMemoryArena::Allocation* alloc, Sizes... sizes) {
	static_assert(sizeof...(Types) == sizeof...(Sizes), "one size per type");
	const size_t s[] = {(size_t)sizes...};
	return distribute_impl<Types...>(alloc, s, std::index_sequence_for<Types...>{});
}
// tcnn::reduce_sum(in, f, out, n, stream): *out += sum f(in[i]) — the device reduction's order is tcnn's own; here one double accumulator
inline size_t reduce_sum_workspace_size(uint32_t) { return 1; }
template <typename T, typename F> void reduce_sum(const T* in, F f, float* out, uint32_t n, cudaStream_t) {
#if defined(__CUDACC__)
	(void)in; (void)f; (void)out; (void)n;
	throw std::runtime_error("reduce_sum: the occupancy update is pinned on the CPU build only");
#else
	double acc = 0.0;
	for (uint32_t i = 0; i < n; ++i) acc += (double)f(in[i]);
	*out += (float)acc;
#endif
}
}  // namespace tcnn

NGP_NAMESPACE_BEGIN
using namespace Eigen;
using namespace tcnn;

struct ScopeGuard {
	std::function<void()> f;
	ScopeGuard(std::function<void()> f_) : f(std::move(f_)) {}
	~ScopeGuard() { if (f) f(); }
};

// nerf_network.h:87-120 — the members the render path calls. The arithmetic is supplied by the test.
#if defined(__CUDACC__)
// = nsb_inference of include/nerfshop_b200.h (device pointers, stream-ordered)
typedef int (*ref_inference_fn)(void* user, const float* coords_dev, uint32_t n, uint16_t* out_dev, uint32_t n_padded, void* stream);
#else
typedef void (*ref_inference_fn)(void* user, const float* coords /*7 x n, column-major*/, uint32_t n, uint16_t* out /*16 x n row-major fp16*/);
#endif
typedef void (*ref_density_fn)(void* user, const float* positions /*3 x n, column-major, warped*/, uint32_t n, uint16_t* out /*16 x n row-major fp16, row 0 = raw density*/);
template <typename T> struct NerfNetwork {
	ref_inference_fn fn = nullptr;
	ref_density_fn dfn = nullptr;  // NerfNetwork::density (nerf_network.h): the occupancy update's call (testbed_nerf.cu:3601-3603)
	void* user = nullptr;
	uint32_t padded_density_output_width() const { return 16; }
	void density(cudaStream_t, const GPUMatrixDynamic<float>& input, GPUMatrixDynamic<T>& output, bool = true) {
		if (!dfn) throw std::runtime_error("NerfNetwork::density call-back not set");
		dfn(user, input.data(), input.n(), (uint16_t*)output.data());
		n_inferred += input.n();
		++n_calls;
	}
	uint64_t n_inferred = 0;
	uint32_t n_calls = 0;
	uint32_t padded_output_width() const { return 16; }
	uint32_t n_extra_dims() const { return 0; }
	void inference_mixed_precision(cudaStream_t stream, const GPUMatrixDynamic<float>& input, GPUMatrixDynamic<T>& output, bool = true) {
		static_assert(sizeof(T) == 2, "network_precision_t is fp16");
#if defined(__CUDACC__)
		if (fn(user, input.data(), input.n(), (uint16_t*)output.data(), output.n(), (void*)stream) != 0) throw std::runtime_error("inference call-back failed");
#else
		if (output.layout() == CM) {  // GrowingSelection::compute_poisson_boundary reads a column-major output (sample-major, 16 per sample)
			std::vector<uint16_t> tmp((size_t)16 * input.n());
			fn(user, input.data(), input.n(), tmp.data());
			uint16_t* o = (uint16_t*)output.data();
			for (uint32_t i = 0; i < input.n(); ++i) for (uint32_t r = 0; r < 16; ++r) o[(size_t)i * 16 + r] = tmp[(size_t)r * input.n() + i];
		} else {
			fn(user, input.data(), input.n(), (uint16_t*)output.data());
		}
#endif
		n_inferred += input.n();
		++n_calls;
	}
	void inference(cudaStream_t, const GPUMatrixDynamic<float>&, GPUMatrixDynamic<float>&) { throw std::runtime_error("NerfNetwork::inference (Slice mode): not on the pinned path"); }
	void input_gradient(cudaStream_t, uint32_t, const GPUMatrixDynamic<float>&, GPUMatrixDynamic<float>&) { throw std::runtime_error("NerfNetwork::input_gradient (Normals mode): needs tiny-cuda-nn"); }
	void visualize_activation(cudaStream_t, uint32_t, uint32_t, const GPUMatrixDynamic<float>&, GPUMatrixDynamic<float>&) { throw std::runtime_error("visualize_activation: needs tiny-cuda-nn"); }
};

// render_buffer.h:189-195
struct CudaRenderBuffer {
	Vector2i res;
	uint32_t m_spp = 0;
	Array4f* fb = nullptr;
	float* depth = nullptr;
	uint32_t spp() const { return m_spp; }
	const Vector2i& in_resolution() const { return res; }
	Array4f* frame_buffer() const { return fb; }
	float* depth_buffer() const { return depth; }
	bool dlss() const { return false; }
};

// m_envmap.envmap / m_distortion.map are tcnn trainable buffers in the reference; here a view of the caller's texels (null = none)
template <typename T> struct NullParams { T* p = nullptr; T* params_inference() const { return p; } };

// testbed.h: the Testbed members read by render_nerf (testbed_nerf.cu:3066-3201) and the NerfTracer class (:129-240), same names
struct Testbed {
	class NerfTracer {
	public:
		NerfTracer() : m_hit_counter(1), m_alive_counter(1) {}
		void init_rays_from_camera(uint32_t spp, uint32_t padded_output_width, uint32_t n_extra_dims, const Eigen::Vector2i& resolution,
		                           const Eigen::Vector2f& focal_length, const Eigen::Matrix<float, 3, 4>& camera_matrix0, const Eigen::Matrix<float, 3, 4>& camera_matrix1,
		                           const Eigen::Vector4f& rolling_shutter, Eigen::Vector2f screen_center, bool snap_to_pixel_centers, const BoundingBox& render_aabb,
		                           float plane_z, float dof, const CameraDistortion& camera_distortion, const float* envmap_data, const Eigen::Vector2i& envmap_resolution,
		                           const float* distortion_data, const Eigen::Vector2i& distortion_resolution, Eigen::Array4f* frame_buffer, float* depth_buffer, uint8_t* grid,
		                           int show_accel, float cone_angle_constant, ERenderMode render_mode, cudaStream_t stream);
		uint32_t trace(NerfNetwork<precision_t>& network, const BoundingBox& render_aabb, const BoundingBox& train_aabb, const uint32_t n_training_images,
		               const TrainingXForm* training_xforms, const Eigen::Vector2f& focal_length, float cone_angle_constant, const uint8_t* grid, ERenderMode render_mode,
		               const Eigen::Matrix<float, 3, 4>& camera_matrix, float depth_scale, int visualized_layer, int visualized_dim, ENerfActivation rgb_activation,
		               ENerfActivation density_activation, int show_accel, float min_transmittance, float glow_y_cutoff, int glow_mode, const Eigen::Vector3f& light_dir,
		               bool apply_operators, cudaStream_t stream);
		void enlarge(size_t n_elements, uint32_t padded_output_width, uint32_t n_extra_dims, cudaStream_t stream);
		RaysNerfSoa& rays_hit() { return m_rays_hit; }
		RaysNerfSoa& rays_init() { return m_rays[0]; }
		uint32_t n_rays_initialized() const { return m_n_rays_initialized; }
		void add_edit_operator(std::shared_ptr<EditOperator> edit_operator) { m_active_edit_operator = m_edit_operators.size(); m_edit_operators.push_back(edit_operator); }
		void reset_edit_operators() { m_edit_operators.clear(); m_active_edit_operator = -1; }
		std::vector<std::shared_ptr<EditOperator>>& edit_operators() { return m_edit_operators; }
		void clear() { m_scratch_alloc = {}; }
		bool m_poisson_target = true;

	private:
		RaysNerfSoa m_rays[2];
		RaysNerfSoa m_rays_hit;
		precision_t* m_network_output;
		precision_t* m_network_output_old;
		float* m_network_input;
		float* m_network_gradient;
		SH9RGB* m_sh_boundary;
		float* m_density_out_boundary;
		float* m_density_residual_boundary;
		tcnn::GPUMemory<uint32_t> m_hit_counter;
		tcnn::GPUMemory<uint32_t> m_alive_counter;
		uint32_t m_n_rays_initialized = 0;
		tcnn::GPUMemoryArena::Allocation m_scratch_alloc;
		std::vector<std::shared_ptr<EditOperator>> m_edit_operators;
		int m_active_edit_operator = -1;
	};

	struct Nerf {
		NerfTracer tracer;
		bool render_with_camera_distortion = false;  // testbed.h:648
		CameraDistortion render_distortion;
		tcnn::GPUMemory<uint8_t> density_grid_bitfield;
		// the occupancy update (testbed.h:620-640)
		tcnn::GPUMemory<float> density_grid, density_grid_mean;
		uint32_t density_grid_ema_step = 0;
		uint32_t max_cascade = 0;
		uint8_t* get_density_grid_bitfield_mip(uint32_t mip) { return density_grid_bitfield.data() + grid_mip_offset(mip) / 8; }
		int show_accel = -1;
		float cone_angle_constant = 1.f / 256.f;
		struct Training {
			struct Dataset { float scale = 0.33f; } dataset;
			uint32_t n_images_for_training = 0;
			tcnn::GPUMemory<TrainingXForm> transforms;
			bool linear_colors = false;
			float density_grid_decay = 0.95f;
		} training;
		ENerfActivation rgb_activation = ENerfActivation::Logistic;
		ENerfActivation density_activation = ENerfActivation::Exponential;
		float rendering_min_transmittance = 0.01f;
		float m_glow_y_cutoff = 0.f;
		int m_glow_mode = 0;
		Eigen::Vector3f light_dir = Eigen::Vector3f::Constant(0.5f);
		tcnn::GPUMemory<float> vis_input;
		tcnn::GPUMemory<Array4f> vis_rgba;
	} m_nerf;
	float m_slice_plane_z = 0.f, m_scale = 1.f;
	ERenderMode m_render_mode = ERenderMode::Shade;
	int m_visualized_dimension = -1, m_visualized_layer = 0;
	bool m_snap_to_pixel_centers = false;
	BoundingBox m_render_aabb, m_aabb;
	float m_dof = 0.f;
	struct { std::shared_ptr<NullParams<float>> envmap = std::make_shared<NullParams<float>>(); Vector2i resolution = Vector2i::Zero(); } m_envmap;
	struct { std::shared_ptr<NullParams<float>> map = std::make_shared<NullParams<float>>(); Vector2i resolution = Vector2i::Zero(); } m_distortion;

	void render_nerf(NerfNetwork<network_precision_t>& network, CudaRenderBuffer& render_buffer, const Vector2i& max_res, const Vector2f& focal_length,
	                 const Matrix<float, 3, 4>& camera_matrix0, const Matrix<float, 3, 4>& camera_matrix1, const Vector4f& rolling_shutter, const Vector2f& screen_center,
	                 bool apply_operators, cudaStream_t stream);
	// the occupancy update through the operators (testbed_nerf.cu:3533-3658)
	std::shared_ptr<NerfNetwork<network_precision_t>> m_nerf_network;
	tcnn::default_rng_t m_rng;
	cudaStream_t m_inference_stream = nullptr;
	void update_density_grid_nerf_operator(uint32_t n_uniform_density_grid_samples, uint32_t n_nonuniform_density_grid_samples, bool reset_grid, cudaStream_t stream);
	void update_density_grid_mean_and_bitfield(cudaStream_t stream);
};

// NsbFrame (ABI 3) -> the Testbed members render_nerf reads for the general camera, glow, environment and distortion maps
static inline void set_frame_extras(Testbed& tb, const NsbFrame* f) {
	tb.m_nerf.render_with_camera_distortion = f->camera_distortion_mode != 0 || f->distortion_dev != nullptr;
	tb.m_nerf.render_distortion.mode = (ECameraDistortionMode)f->camera_distortion_mode;
	for (int i = 0; i < 7; ++i) tb.m_nerf.render_distortion.params[i] = f->camera_distortion_params[i];
	tb.m_dof = f->dof;
	tb.m_slice_plane_z = 0.f;
	tb.m_scale = f->focus_z > 0.f ? f->focus_z : 1.f;  // plane_z = m_slice_plane_z + m_scale
	tb.m_nerf.m_glow_mode = f->glow_mode;
	tb.m_nerf.m_glow_y_cutoff = f->glow_y_cutoff;
	tb.m_envmap.envmap->p = const_cast<float*>(f->envmap_dev);
	tb.m_envmap.resolution = Vector2i(f->envmap_resolution[0], f->envmap_resolution[1]);
	tb.m_distortion.map->p = const_cast<float*>(f->distortion_dev);
	tb.m_distortion.resolution = Vector2i(f->distortion_resolution[0], f->distortion_resolution[1]);
}
static inline CameraDistortion frame_camera_distortion(const NsbFrame* f) {
	CameraDistortion cd;
	cd.mode = (ECameraDistortionMode)f->camera_distortion_mode;
	for (int i = 0; i < 7; ++i) cd.params[i] = f->camera_distortion_params[i];
	return cd;
}

// ---- the reference's code ---------------------------------------------------------------------------------------------
#include "testbed_nerf.inc"
#ifndef __CUDACC__  // frame post-process: CPU build only (the tonemap kernel writes a GL surface on the GPU; here only its `lopi` copy is kept)
typedef unsigned long long cudaSurfaceObject_t;
template <typename V> inline void surf2Dwrite(V, cudaSurfaceObject_t, size_t, uint32_t) {}
#include "render_buffer.inc"
// growing_selection.h's members that compute_poisson_boundary touches (growing_selection.cu:2220-2348), same names
SH9RGB project_sh9(const Eigen::Vector3f& dir, const Eigen::Vector3f& rgb, const float domega = 1.0f);
struct GrowingSelection {
	using point_t = Eigen::Vector3f;
	Cage<float, point_t> proxy_cage;
	struct { uint32_t sh_sampling_width = 10; } m_poisson_editing;
	uint32_t m_hemisphere_width = 10;
	std::shared_ptr<NerfNetwork<network_precision_t>> m_nerf_network;
	BoundingBox m_aabb;
	cudaStream_t m_stream = nullptr;
	ENerfActivation m_rgb_activation = ENerfActivation::Logistic, m_density_activation = ENerfActivation::Exponential;
	tcnn::GPUMemory<uint8_t> m_density_grid_bitfield;
	void compute_poisson_boundary(const bool is_inside);
};
#include "sh_utils.inc"
#include "growing_selection_boundary.inc"
#endif
#include "cage_deformation.inc"
#include "affine_duplication.inc"
#include "selection_utils.inc"
#ifndef __CUDACC__  // host-side per-edit geometry: CPU build only
#include "tet_mesh.inc"
#include "cage.inc"
#endif


// ---------------------------------------------------------------------------------------------------------------------
// EditOperator subclasses: the argument marshalling of CageDeformation::{map_rays, compute_poisson_full_residuals}
// (cage_deformation.cu:547-575, 675-716) and AffineDuplication::map_rays (affine_duplication.cu:120-135) from the C-ABI POD
// ---------------------------------------------------------------------------------------------------------------------
static BoundingBox bb(const float* mn, const float* mx) { return BoundingBox(Vector3f(mn[0], mn[1], mn[2]), Vector3f(mx[0], mx[1], mx[2])); }

struct RefOpBase : EditOperator {
	bool imgui(bool&, const Vector2i&, const Vector2f&, const Matrix<float, 3, 4>&, const Vector2f&, bool&) override { return false; }
	bool visualize_edit_gui(const Matrix<float, 4, 4>&, const Matrix<float, 4, 4>&, const Matrix<float, 4, 4>&, const Vector2f&, float, float) override { return false; }
	void draw_gl(const Vector2i&, const Vector2f&, const Matrix<float, 3, 4>&, const Vector2f&) override {}
	bool handle_keyboard() override { return false; }
	nlohmann::json to_json() override { return {}; }
};

struct RefCageOp : RefOpBase {
	NsbEditOp op;
	explicit RefCageOp(const NsbEditOp& o) : op(o) { static_assert(sizeof(SH9RGB) == 27 * sizeof(float), "SH9RGB = 9 x 3 floats, column-major, like the C-ABI array"); }
	void map_rays(cudaStream_t stream, tcnn::PitchedPtr<NerfCoordinate> nerf_coords, tcnn::GPUMatrixDynamic<bool>& empty_mask, uint32_t n_elements) const override {
		if (op.n_tets == 0) return;
		tcnn::linear_kernel(interpolate_tet, 0, stream, n_elements, nerf_coords, empty_mask.data(), (bool)op.copy, bb(op.scene_aabb_min, op.scene_aabb_max),
		                    bb(op.warped_bbox_min, op.warped_bbox_max), bb(op.original_warped_bbox_min, op.original_warped_bbox_max), op.tet_lut_idx, op.tet_lut_offsets, op.tets,
		                    (const Vector3f*)op.vertices, (const Vector3f*)op.original_vertices, (const Matrix3f*)op.local_rotations, op.original_bitfield);
	}
	// CageDeformation::map_positions (cage_deformation.cu:624-645) and ::compute_poisson_residual_density (:647-672); no cutting plane (m_plane_dir = 0)
	void map_positions(cudaStream_t stream, tcnn::PitchedPtr<NerfPosition> nerf_pos, tcnn::GPUMatrixDynamic<bool>& empty_mask, uint32_t n_elements) const override {
		if (op.n_tets == 0) return;
		tcnn::linear_kernel(interpolate_tet_pos, 0, stream, n_elements, nerf_pos, empty_mask.data(), bb(op.scene_aabb_min, op.scene_aabb_max), bb(op.warped_bbox_min, op.warped_bbox_max),
		                    bb(op.original_warped_bbox_min, op.original_warped_bbox_max), op.tet_lut_idx, op.tet_lut_offsets, op.tets, (const Vector3f*)op.vertices,
		                    (const Vector3f*)op.original_vertices, op.original_bitfield);
	}
	void compute_poisson_residual_density(cudaStream_t stream, const uint32_t n_elements, tcnn::PitchedPtr<NerfPosition> input_position,
	                                      tcnn::network_precision_t* density_network_output) const override {
		if (!op.apply_poisson || op.n_tets == 0 || !op.boundary_residual_density) return;
		tcnn::linear_kernel(compute_poisson_residual_density_kernel, 0, stream, n_elements, input_position, density_network_output, bb(op.scene_aabb_min, op.scene_aabb_max),
		                    Vector3f(Vector3f::Zero()), Vector3f(Vector3f::Zero()), bb(op.bbox_min, op.bbox_max), op.tet_lut_idx, op.tet_lut_offsets, op.tets,
		                    (const Vector3f*)op.vertices, op.boundary_residual_density);
	}
	void compute_poisson_full_residuals(cudaStream_t stream, const uint32_t n_elements, NerfPayload* payloads, tcnn::PitchedPtr<NerfCoordinate> network_input,
	                                    SH9RGB* __restrict__ sh_boundary, float* __restrict__ out_density_boundary, float* __restrict__ residual_density_boundary) const override {
		if (!op.apply_poisson || op.n_tets == 0 || !op.boundary_shs || !op.boundary_outside_density || !op.boundary_residual_density) return;
		tcnn::linear_kernel(compute_residual_poisson_kernel, 0, stream, n_elements, payloads, network_input, sh_boundary, out_density_boundary, residual_density_boundary,
		                    bb(op.scene_aabb_min, op.scene_aabb_max), BoundingBox(), bb(op.bbox_min, op.bbox_max), (const uint32_t*)nullptr, op.tet_lut_idx,
		                    (const uint32_t*)nullptr, op.tet_lut_offsets, op.tets, (const Vector3f*)op.original_vertices, (const Vector3f*)op.vertices, (const SH9RGB*)op.boundary_shs,
		                    op.boundary_outside_density, op.boundary_residual_density, op.residual_amplitude);
	}
};

static AffineBoundingBox abox(const NsbAffineBox& b) {
	AffineBoundingBox r;
	r.min = Vector3f(b.min[0], b.min[1], b.min[2]);
	r.u = Vector3f(b.u[0], b.u[1], b.u[2]);
	r.v = Vector3f(b.v[0], b.v[1], b.v[2]);
	r.w = Vector3f(b.w[0], b.w[1], b.w[2]);
	r.center = Vector3f(b.center[0], b.center[1], b.center[2]);
	r.max = r.min; r.scale = Vector3f::Ones();
	return r;
}
struct RefAffineOp : RefOpBase {
	NsbEditOp op;
	explicit RefAffineOp(const NsbEditOp& o) : op(o) {}
	void map_rays(cudaStream_t stream, tcnn::PitchedPtr<NerfCoordinate> nerf_coords, tcnn::GPUMatrixDynamic<bool>& empty_mask, uint32_t n_elements) const override {
		Matrix3f R; memcpy(R.data(), op.rotation, 9 * sizeof(float));
		tcnn::linear_kernel(translate_in_box, 0, stream, n_elements, nerf_coords, op.hide_original ? empty_mask.data() : (bool*)nullptr, abox(op.selection_box),
		                    abox(op.destination_box), Vector3f(op.translation[0], op.translation[1], op.translation[2]), Vector3f(op.scale[0], op.scale[1], op.scale[2]), R,
		                    (bool)op.correct_dir);
	}
	// AffineDuplication::map_positions (affine_duplication.cu:137-150)
	void map_positions(cudaStream_t stream, tcnn::PitchedPtr<NerfPosition> nerf_pos, tcnn::GPUMatrixDynamic<bool>& empty_mask, uint32_t n_elements) const override {
		Matrix3f R; memcpy(R.data(), op.rotation, 9 * sizeof(float));
		tcnn::linear_kernel(translate_in_box_pos, 0, stream, n_elements, nerf_pos, op.hide_original ? empty_mask.data() : (bool*)nullptr, abox(op.selection_box), abox(op.destination_box),
		                    Vector3f(op.translation[0], op.translation[1], op.translation[2]), Vector3f(op.scale[0], op.scale[1], op.scale[2]), R);
	}
};
static std::shared_ptr<EditOperator> make_op(const NsbEditOp& o) {
	if (o.type == NSB_OP_CAGE) return std::make_shared<RefCageOp>(o);
	return std::make_shared<RefAffineOp>(o);
}

NGP_NAMESPACE_END
