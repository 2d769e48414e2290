// tiny-cuda-nn/common.h — TEST INFRASTRUCTURE ONLY (oracle/_ref). tiny-cuda-nn is an absent submodule of the reference
// (SURVEY.md §0); this restates the handful of public helpers of that library that the reference's render-path sources call
// (names and semantics of NVlabs/tiny-cuda-nn common.h / common_device.h of mid-2022).
#pragma once
#include "../cuda_on_cpu.h"

#define TCNN_NAMESPACE_BEGIN namespace tcnn {
#define TCNN_NAMESPACE_END }
#if defined(__CUDACC__)
#define TCNN_HOST_DEVICE __host__ __device__
#else
#define TCNN_HOST_DEVICE
#endif
#define TCNN_MIN_GPU_ARCH 0
#define TCNN_PRAGMA_UNROLL

namespace tcnn {
using network_precision_t = __half;
static constexpr uint32_t batch_size_granularity = 128;
static constexpr uint32_t n_threads_linear = 128;
enum MatrixLayout { RowMajor = 0, SoA = 0, ColumnMajor = 1, AoS = 1 };
static constexpr MatrixLayout RM = RowMajor, CM = ColumnMajor;

template <typename T> TCNN_HOST_DEVICE inline T div_round_up(T a, T b) { return (a + b - 1) / b; }
template <typename T> TCNN_HOST_DEVICE inline T next_multiple(T a, T b) { return div_round_up(a, b) * b; }
template <typename T> TCNN_HOST_DEVICE inline T clamp(T v, T lo, T hi) { return v < lo ? lo : (hi < v ? hi : v); }  // min(max(val, lower), upper)
template <typename T> TCNN_HOST_DEVICE inline void host_device_swap(T& a, T& b) { T c(a); a = b; b = c; }
TCNN_HOST_DEVICE inline float logistic(const float x) { return 1.0f / (1.0f + expf(-x)); }
TCNN_HOST_DEVICE inline float logit(const float x) { return -logf(1.0f / (fminf(fmaxf(x, 1e-9f), 1.0f - 1e-9f)) - 1.0f); }

TCNN_HOST_DEVICE inline uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
TCNN_HOST_DEVICE inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
TCNN_HOST_DEVICE inline uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249;
	x = (x | (x >> 2)) & 0xc30c30c3;
	x = (x | (x >> 4)) & 0x0f00f00f;
	x = (x | (x >> 8)) & 0xff0000ff;
	x = (x | (x >> 16)) & 0x0000ffff;
	return x;
}

template <typename T, uint32_t N> struct alignas(sizeof(T) * N > 16 ? 16 : sizeof(T) * N) vector_t {
	T data[N];
	TCNN_HOST_DEVICE T& operator[](uint32_t i) { return data[i]; }
	TCNN_HOST_DEVICE const T& operator[](uint32_t i) const { return data[i]; }
	static constexpr uint32_t N_ELEMS = N;
};

template <typename T> struct PitchedPtr {
	TCNN_HOST_DEVICE PitchedPtr() : ptr{nullptr}, stride_in_bytes{sizeof(T)} {}
	TCNN_HOST_DEVICE PitchedPtr(T* ptr, size_t stride_in_elements, size_t offset = 0, size_t extra_stride_bytes = 0)
	    : ptr{ptr + offset}, stride_in_bytes{(uint32_t)(stride_in_elements * sizeof(T) + extra_stride_bytes)} {}
	template <typename U> TCNN_HOST_DEVICE explicit PitchedPtr(PitchedPtr<U> other) : ptr{(T*)other.ptr}, stride_in_bytes{other.stride_in_bytes} {}
	TCNN_HOST_DEVICE T* operator()(uint32_t y) const { return (T*)((const char*)ptr + y * stride_in_bytes); }
	TCNN_HOST_DEVICE void operator+=(uint32_t y) { ptr = (T*)((const char*)ptr + y * stride_in_bytes); }
	TCNN_HOST_DEVICE void operator-=(uint32_t y) { ptr = (T*)((const char*)ptr - y * stride_in_bytes); }
	TCNN_HOST_DEVICE explicit operator bool() const { return ptr; }
	T* ptr;
	uint32_t stride_in_bytes;
};

#if defined(__CUDACC__)
template <typename K, typename T, typename... Types> inline void linear_kernel(K kernel, uint32_t shmem_size, cudaStream_t stream, T n_elements, Types... args) {
	if (n_elements <= 0) return;
	kernel<<<div_round_up((uint32_t)n_elements, n_threads_linear), n_threads_linear, shmem_size, stream>>>((uint32_t)n_elements, args...);
}
#else
// A kernel launch = a loop over the grid. ref_driver.cpp decides whether the loop is parallel (OpenMP) through this hook.
extern bool g_parallel_launch;
template <typename K, typename T, typename... Types> inline void linear_kernel(K kernel, uint32_t shmem_size, cudaStream_t stream, T n_elements, Types... args) {
	if (n_elements <= 0) return;
	const uint32_t n = (uint32_t)n_elements;
	const uint32_t blocks = div_round_up(n, n_threads_linear);
	if (g_parallel_launch) {
#pragma omp parallel for schedule(dynamic, 4)
		for (uint32_t b = 0; b < blocks; ++b) {
			blockDim = dim3(n_threads_linear); gridDim = dim3(blocks); blockIdx = {b, 0, 0};
			for (uint32_t t = 0; t < n_threads_linear; ++t) { threadIdx = {t, 0, 0}; kernel((uint32_t)n_elements, args...); }
		}
	} else {
		for (uint32_t b = 0; b < blocks; ++b) {
			blockDim = dim3(n_threads_linear); gridDim = dim3(blocks); blockIdx = {b, 0, 0};
			for (uint32_t t = 0; t < n_threads_linear; ++t) { threadIdx = {t, 0, 0}; kernel((uint32_t)n_elements, args...); }
		}
	}
}
#endif
}  // namespace tcnn
