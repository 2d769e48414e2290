// nerfshop_host.hpp — C++ shim over the C ABI (include/nerfshop_b200.h) that keeps the reference's class names and
// member signatures for the render path, so that Testbed::render_frame (src/testbed.cu:2632-2648) can call it the way
// it calls the stock code. Header-only; Eigen-free (cameras are passed as 12 floats, column-major like Eigen's storage).
//
//   ngp_b200::NerfNetwork      <- NerfNetwork<T>            (nerf_network.h:87-120): inference_mixed_precision, density
//   ngp_b200::EditOperator     <- EditOperator              (editing/edit_operator.h:25-94): the POD handed to the tracer
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
// stays in the host application. CageDeformation / AffineDuplication fill the respective half of NsbEditOp.
struct EditOperator {
	NsbEditOp pod{};
	virtual ~EditOperator() = default;
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

	// init_rays_from_camera + trace + shade in one call; returns n_hit like NerfTracer::trace.
	uint32_t trace(const NsbFrame& frame, float* frame_buffer_dev, float* depth_buffer_dev, void* stream) {
		NsbFrame f = frame;
		f.poisson_target = m_poisson_target ? 1 : 0;
		check(nsb_render(m_ctx->get(), &f, frame_buffer_dev, depth_buffer_dev, stream), "nsb_render");
		m_rendered = true;
		return 0;  // asynchronous: n_hit is available from stats() (the reference synchronises here, :2998-3000)
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
		for (auto& op : m_edit_operators) pods.push_back(op->pod);
		check(nsb_set_edit_ops(m_ctx->get(), pods.empty() ? nullptr : pods.data(), (int32_t)pods.size()), "nsb_set_edit_ops");
	}
	std::shared_ptr<Context> m_ctx;
	std::vector<std::shared_ptr<EditOperator>> m_edit_operators;
	bool m_rendered = false;
};

// Testbed::render_nerf: the caller fills NsbFrame from the Testbed members listed at testbed_nerf.cu:3082-3138.
inline void render_nerf(NerfTracer& tracer, const NsbFrame& frame, float* frame_buffer_dev, float* depth_buffer_dev, bool apply_operators, void* stream) {
	NsbFrame f = frame;
	f.apply_operators = apply_operators ? 1 : 0;
	tracer.trace(f, frame_buffer_dev, depth_buffer_dev, stream);
}

}  // namespace ngp_b200
