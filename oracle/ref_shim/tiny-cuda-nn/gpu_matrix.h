// tiny-cuda-nn/gpu_matrix.h — TEST INFRASTRUCTURE ONLY (oracle/_ref): host-memory stand-in for the GPUMatrix views/owners.
#pragma once
#include <memory>
#include <tiny-cuda-nn/gpu_memory.h>
namespace tcnn {
template <typename T> struct GPUMatrixDynamic {
	T* m_data = nullptr; uint32_t m_rows = 0, m_cols = 0; MatrixLayout m_layout = CM; std::shared_ptr<void> m_owned;
	GPUMatrixDynamic() = default;
	GPUMatrixDynamic(T* data, uint32_t m, uint32_t n, MatrixLayout layout = CM) : m_data(data), m_rows(m), m_cols(n), m_layout(layout) {}
	GPUMatrixDynamic(uint32_t m, uint32_t n, MatrixLayout layout = CM) : m_rows(m), m_cols(n), m_layout(layout) {
#if defined(__CUDACC__)
		void* p = nullptr; CUDA_CHECK_THROW(cudaMalloc(&p, (size_t)m * n * sizeof(T) + 16)); m_owned = std::shared_ptr<void>(p, [](void* q) { cudaFree(q); });
#else
		m_owned = std::shared_ptr<void>(calloc((size_t)m * n * sizeof(T) + 16, 1), free);
#endif
		m_data = (T*)m_owned.get();
	}
	T* data() const { return m_data; }
	uint32_t rows() const { return m_rows; } uint32_t m() const { return m_rows; }
	uint32_t cols() const { return m_cols; } uint32_t n() const { return m_cols; }
	size_t n_elements() const { return (size_t)m_rows * m_cols; }
	size_t n_bytes() const { return n_elements() * sizeof(T); }
	MatrixLayout layout() const { return m_layout; }
};
template <typename T, MatrixLayout L = CM> struct GPUMatrix : GPUMatrixDynamic<T> {
	GPUMatrix() = default;
	GPUMatrix(T* data, uint32_t m, uint32_t n) : GPUMatrixDynamic<T>(data, m, n, L) {}
	GPUMatrix(uint32_t m, uint32_t n) : GPUMatrixDynamic<T>(m, n, L) {}
};
}  // namespace tcnn
