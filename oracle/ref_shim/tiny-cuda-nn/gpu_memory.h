// tiny-cuda-nn/gpu_memory.h — TEST INFRASTRUCTURE ONLY (oracle/_ref): host-memory stand-in for tcnn::GPUMemory<T>.
#pragma once
#include <tiny-cuda-nn/common.h>
#if defined(__CUDACC__)
namespace tcnn {
template <typename T> struct GPUMemory {
	T* m_data = nullptr; size_t m_size = 0;
	GPUMemory() = default;
	explicit GPUMemory(size_t n) { resize(n); }
	GPUMemory(const GPUMemory&) = delete;
	GPUMemory& operator=(const GPUMemory&) = delete;
	GPUMemory(GPUMemory&& o) noexcept : m_data(o.m_data), m_size(o.m_size) { o.m_data = nullptr; o.m_size = 0; }
	GPUMemory& operator=(GPUMemory&& o) noexcept { std::swap(m_data, o.m_data); std::swap(m_size, o.m_size); return *this; }
	~GPUMemory() { if (m_data) cudaFree(m_data); }
	void resize(size_t n) { if (n == m_size) return; if (m_data) { cudaFree(m_data); m_data = nullptr; } m_size = n; if (n) CUDA_CHECK_THROW(cudaMalloc((void**)&m_data, n * sizeof(T))); }
	void enlarge(size_t n) { if (n > m_size) resize(n); }
	void memset(int value) { if (m_size) CUDA_CHECK_THROW(cudaMemset(m_data, value, m_size * sizeof(T))); }
	void free_memory() { resize(0); }
	void copy_from_host(const std::vector<T>& h) { copy_from_host(h.data(), h.size()); }
	void copy_from_host(const T* h, size_t n) { if (n > m_size) resize(n); if (n) CUDA_CHECK_THROW(cudaMemcpy(m_data, h, n * sizeof(T), cudaMemcpyHostToDevice)); }
	void copy_from_host(const T* h) { copy_from_host(h, m_size); }
	void resize_and_copy_from_host(const std::vector<T>& h) { resize(h.size()); copy_from_host(h); }
	void copy_to_host(std::vector<T>& h) const { h.resize(m_size); copy_to_host(h.data(), m_size); }
	void copy_to_host(T* h, size_t n) const { if (n) CUDA_CHECK_THROW(cudaMemcpy(h, m_data, n * sizeof(T), cudaMemcpyDeviceToHost)); }
	T* data() const { return m_data; }
	size_t size() const { return m_size; }
	size_t get_num_elements() const { return m_size; }
	size_t n_bytes() const { return m_size * sizeof(T); }
	size_t get_bytes() const { return m_size * sizeof(T); }
	size_t bytes() const { return m_size * sizeof(T); }
};
}  // namespace tcnn
#else
namespace tcnn {
template <typename T> struct GPUMemory {
	std::vector<T> v;
	GPUMemory() = default;
	explicit GPUMemory(size_t n) : v(n) {}
	void resize(size_t n) { v.resize(n); }
	void enlarge(size_t n) { if (n > v.size()) v.resize(n); }
	void memset(int value) { if (!v.empty()) ::memset((void*)v.data(), value, v.size() * sizeof(T)); }
	void free_memory() { v.clear(); v.shrink_to_fit(); }
	void copy_from_host(const std::vector<T>& h) { v = h; }
	void copy_from_host(const T* h, size_t n) { v.assign(h, h + n); }
	void copy_from_host(const T* h) { std::copy(h, h + v.size(), v.begin()); }
	void copy_to_host(std::vector<T>& h) const { h = v; }
	void copy_to_host(T* h, size_t n) const { std::copy(v.begin(), v.begin() + n, h); }
	void resize_and_copy_from_host(const std::vector<T>& h) { v = h; }
	T* data() { return v.data(); }
	const T* data() const { return v.data(); }
	size_t size() const { return v.size(); }
	size_t get_num_elements() const { return v.size(); }
	size_t n_bytes() const { return v.size() * sizeof(T); }
	size_t get_bytes() const { return v.size() * sizeof(T); }
	size_t bytes() const { return v.size() * sizeof(T); }
	T& operator[](size_t i) { return v[i]; }
	const T& operator[](size_t i) const { return v[i]; }
};
}  // namespace tcnn
#endif
