// nerfshop_host.hpp — C++ shim over the C ABI (include/nerfshop_b200.h) that keeps the reference's class names and
// member signatures for the render path, so that Testbed::render_frame (src/testbed.cu:2632-2648) can call it the way
// it calls the stock code. Header-only; Eigen-free (cameras are passed as 12 floats, column-major like Eigen's storage).
//
//   ngp_b200::NerfNetwork      <- NerfNetwork<T>            (nerf_network.h:87-120): inference_mixed_precision, density
//   ngp_b200::EditOperator     <- EditOperator              (editing/edit_operator.h:25-94): the POD handed to the tracer + the four device
//                                                            virtuals (map_rays, map_positions, compute_poisson_full_residuals,
//                                                            compute_poisson_residual_density) once the operator is attached to a tracer
//   ngp_b200::NerfTracer       <- Testbed::NerfTracer       (testbed.h:129-240): operator list management, trace == render
//   ngp_b200::render_nerf(...) <- Testbed::render_nerf      (testbed.h:305, testbed_nerf.cu:3066)
//
// Errors become std::runtime_error, as CUDA_CHECK_THROW does in the reference (propagates to main.cu:252).
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nerfshop_b200.h"

namespace ngp_b200 {

inline void check(NsbStatus s, const char* what) {
	if (s != NSB_OK) throw std::runtime_error(std::string(what) + ": " + nsb_last_error());
}

class Context {
public:
	explicit Context(int device = -1) { check(nsb_create(device, &m_ctx), "nsb_create"); }
	~Context() { nsb_destroy(m_ctx); }
	Context(const Context&) = delete;
	Context& operator=(const Context&) = delete;
	NsbContext* get() const { return m_ctx; }

private:
	NsbContext* m_ctx = nullptr;
};

// The network as the render path sees it. `set_params` takes the contiguous fp16 block tcnn::Trainer owns
// ([density MLP][rgb MLP][hash grid], nerf_network_full.h:316-349).
class NerfNetwork {
public:
	NerfNetwork(std::shared_ptr<Context> ctx, const NsbModelDesc& desc) : m_ctx(std::move(ctx)), m_desc(desc) {}
	void set_params(const uint16_t* params_fp16, uint64_t n_params) { check(nsb_upload_model(m_ctx->get(), &m_desc, params_fp16, n_params), "nsb_upload_model"); }
	// the block where tcnn::Trainer keeps it (device memory): no 26 MB round trip through the host
	void set_params_device(const uint16_t* params_fp16_dev, uint64_t n_params) { check(nsb_upload_model_dev(m_ctx->get(), &m_desc, params_fp16_dev, n_params), "nsb_upload_model_dev"); }
	// tcnn's FullyFusedMLP accumulates in __half fragments; NSB_MLP_ACC_F32 (default) accumulates wide (see include/nerfshop_b200.h)
	void set_accumulator(NsbMlpAccumulator policy) { check(nsb_set_mlp_accumulator(m_ctx->get(), policy), "nsb_set_mlp_accumulator"); }
	uint64_t n_params() const { uint64_t n = 0; check(nsb_model_n_params(&m_desc, &n), "nsb_model_n_params"); return n; }
	uint32_t padded_output_width() const { return 16; }          // nerf_network.h: rgb MLP output padded to 16
	uint32_t padded_density_output_width() const { return 16; }
	uint32_t n_extra_dims() const { return 0; }
	// input: 7 x n floats column-major (NerfCoordinate per column), output: fp16 row-major [16 x n_padded] (testbed_nerf.cu:2890-2913)
	void inference_mixed_precision(void* stream, const float* input_dev, uint16_t* output_dev, uint32_t n, uint32_t n_padded) const {
		check(nsb_inference(m_ctx->get(), input_dev, n, output_dev, n_padded, stream), "nsb_inference");
	}
	void density(void* stream, const float* input_dev, uint16_t* output_dev, uint32_t n, uint32_t n_padded) const {
		check(nsb_density(m_ctx->get(), input_dev, n, output_dev, n_padded, stream), "nsb_density");
	}

private:
	std::shared_ptr<Context> m_ctx;
	NsbModelDesc m_desc;
};

// An operator is the POD of its kernel arguments; the interactive tooling that produces it (GrowingSelection, TetGen, ...)
// stays in the host application. CageDeformation / AffineDuplication fill the respective half of NsbEditOp. Once added to a
// NerfTracer the operator knows its slot in the uploaded list and the device virtuals of the reference's EditOperator work
// (edit_operator.h:43,45,68,81; PitchedPtr arguments become pointer + stride, GPUMatrixDynamic<bool> a byte array).
struct EditOperator {
	NsbEditOp pod{};
	virtual ~EditOperator() = default;
	// nerf_coords: NerfCoordinate per sample (7 floats); empty_mask is only ever set, like interpolate_tet / translate_in_box
	virtual void map_rays(void* stream, float* nerf_coords_dev, uint8_t* empty_mask_dev, uint32_t n_elements) const {
		check(nsb_map_rays_op(ctx(), m_index, nerf_coords_dev, empty_mask_dev, n_elements, stream), "nsb_map_rays_op");
	}
	virtual void map_positions(void* stream, float* nerf_pos_dev, uint32_t stride_floats, uint8_t* empty_mask_dev, uint32_t n_elements) const {
		check(nsb_map_positions(ctx(), m_index, nerf_pos_dev, stride_floats, empty_mask_dev, n_elements, stream), "nsb_map_positions");
	}
	virtual void compute_poisson_full_residuals(void* stream, uint32_t n_elements, const float* network_input_dev, float* sh_boundary_dev, float* out_density_boundary_dev,
	                                            float* residual_density_boundary_dev) const {
		check(nsb_poisson_residuals_op(ctx(), m_index, network_input_dev, n_elements, sh_boundary_dev, out_density_boundary_dev, residual_density_boundary_dev, stream), "nsb_poisson_residuals_op");
	}
	virtual void compute_poisson_residual_density(void* stream, uint32_t n_elements, const float* input_position_dev, uint32_t stride_floats, uint16_t* density_network_output_dev) const {
		check(nsb_poisson_residual_density(ctx(), m_index, input_position_dev, stride_floats, density_network_output_dev, n_elements, stream), "nsb_poisson_residual_density");
	}

private:
	friend class NerfTracer;
	NsbContext* ctx() const {
		if (!m_ctx || m_index < 0) throw std::runtime_error("EditOperator: not attached to a NerfTracer (add_edit_operator first)");
		return m_ctx;
	}
	NsbContext* m_ctx = nullptr;
	int32_t m_index = -1;
};
struct CageDeformation : EditOperator {
	CageDeformation() { pod.type = NSB_OP_CAGE; pod.residual_amplitude = 1.0f; }
};
struct AffineDuplication : EditOperator {
	AffineDuplication() { pod.type = NSB_OP_AFFINE; pod.scale[0] = pod.scale[1] = pod.scale[2] = 1.0f; pod.rotation[0] = pod.rotation[4] = pod.rotation[8] = 1.0f; }
};

class NerfTracer {
public:
	explicit NerfTracer(std::shared_ptr<Context> ctx) : m_ctx(std::move(ctx)) {}
	void add_edit_operator(std::shared_ptr<EditOperator> op) { m_edit_operators.push_back(std::move(op)); upload(); }
	void delete_edit_operator(size_t i) { m_edit_operators.erase(m_edit_operators.begin() + (long)i); upload(); }
	void reset_edit_operator() { m_edit_operators.clear(); upload(); }
	void update_edit_operators() { upload(); }  // after a gizmo drag changed an operator's arrays
	const std::vector<std::shared_ptr<EditOperator>>& edit_operators() const { return m_edit_operators; }
	bool m_poisson_target = false;

	// init_rays_from_camera + trace + shade in one call. Like NerfTracer::trace (:2998-3000) it synchronises the stream and returns n_hit;
	// synchronise = false keeps the frame asynchronous (returns 0; n_hit is in stats() later).
	uint32_t trace(const NsbFrame& frame, float* frame_buffer_dev, float* depth_buffer_dev, void* stream, bool synchronise = true) {
		NsbFrame f = frame;
		f.poisson_target = m_poisson_target ? 1 : 0;
		check(nsb_render(m_ctx->get(), &f, frame_buffer_dev, depth_buffer_dev, stream), "nsb_render");
		m_rendered = true;
		return synchronise ? (uint32_t)stats().n_hit : 0u;
	}
	// ---- per-edit work kept on the device ----
	// GrowingSelection::update_tet_mesh for operator i after a gizmo drag (growing_selection.cu:1615): no re-upload.
	void attach_mvc(size_t i, const float* mvc_weights, uint32_t n_cage_vertices) { check(nsb_cage_attach_mvc(m_ctx->get(), (int32_t)i, mvc_weights, n_cage_vertices), "nsb_cage_attach_mvc"); }
	void deform_cage(size_t i, const float* cage_vertices, uint32_t n_cage_vertices, void* stream) { check(nsb_cage_deform(m_ctx->get(), (int32_t)i, cage_vertices, n_cage_vertices, stream), "nsb_cage_deform"); }
	// GrowingSelection::interpolate_poisson_boundary (growing_selection.cu:2350-2398)
	void set_membrane(size_t i, const float* gamma, uint32_t n_cage_vertices, const float* d_in, const float* d_out, const float* shs_in, const float* shs_out, float amplitude, bool apply) {
		check(nsb_cage_set_membrane(m_ctx->get(), (int32_t)i, gamma, n_cage_vertices, d_in, d_out, shs_in, shs_out, amplitude, apply ? 1 : 0, nullptr, nullptr, nullptr), "nsb_cage_set_membrane");
	}
	// Testbed::update_density_grid_nerf_render (testbed_nerf.cu:3514-3520): the caller owns m_rng / density_grid_ema_step like Testbed does.
	void update_density_grid(const NsbGridUpdate& u, void* stream) { check(nsb_update_density_grid(m_ctx->get(), &u, stream), "nsb_update_density_grid"); }
	NsbRenderStats stats() const { NsbRenderStats s{}; check(nsb_get_stats(m_ctx->get(), &s), "nsb_get_stats"); return s; }
	void clear() {}  // the reference frees 4.6 GB of arena scratch here (:3074-3076); nothing to free

private:
	void upload() {
		std::vector<NsbEditOp> pods;
		for (size_t i = 0; i < m_edit_operators.size(); ++i) {
			m_edit_operators[i]->m_ctx = m_ctx->get();
			m_edit_operators[i]->m_index = (int32_t)i;
			pods.push_back(m_edit_operators[i]->pod);
		}
		check(nsb_set_edit_ops(m_ctx->get(), pods.empty() ? nullptr : pods.data(), (int32_t)pods.size()), "nsb_set_edit_ops");
	}
	std::shared_ptr<Context> m_ctx;
	std::vector<std::shared_ptr<EditOperator>> m_edit_operators;
	bool m_rendered = false;
};

// Testbed::render_nerf: the caller fills NsbFrame from the Testbed members listed at testbed_nerf.cu:3082-3138.
inline uint32_t render_nerf(NerfTracer& tracer, const NsbFrame& frame, float* frame_buffer_dev, float* depth_buffer_dev, bool apply_operators, void* stream, bool synchronise = false) {
	NsbFrame f = frame;
	f.apply_operators = apply_operators ? 1 : 0;
	return tracer.trace(f, frame_buffer_dev, depth_buffer_dev, stream, synchronise);
}

}  // namespace ngp_b200
