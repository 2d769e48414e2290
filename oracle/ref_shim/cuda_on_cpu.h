// cuda_on_cpu.h — TEST INFRASTRUCTURE ONLY (oracle/_ref). Lets the reference's CUDA sources compile for the host with g++:
// qualifiers vanish, a __global__ kernel is an ordinary function, threadIdx/blockIdx/blockDim are thread-local variables that
// the launcher (tcnn::linear_kernel below / ref_driver.cpp) sets before each call, device intrinsics map to libm.
// __expf maps to expf: the hardware's ex2.approx path is not reproducible on a CPU (that difference is inside the 1e-3
// RGBA tolerance and is never part of a bit-exact comparison).
#pragma once
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#if defined(__CUDACC__)
// nvcc build (oracle/ref_driver_cuda.cu): the real CUDA runtime, the reference's kernels run on the GPU
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define CUDA_CHECK_THROW(x)                                                                                         \
	do {                                                                                                            \
		cudaError_t result_ = (x);                                                                                  \
		if (result_ != cudaSuccess) throw std::runtime_error(std::string(#x " failed: ") + cudaGetErrorString(result_)); \
	} while (0)
// kernel<<<blocks, threads, shmem, stream>>>(args...) was rewritten by oracle/ref_build.py into nsb_launch(blocks, threads, kernel, args...)
// (the CPU build cannot parse the launch syntax); here it is the same launch again, on the default stream like the call sites' `stream`
// argument in this harness.
template <typename K, typename... Args> inline void nsb_launch(dim3 blocks, dim3 threads, K kernel, Args... args) { kernel<<<blocks, threads>>>(args...); }
#else
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#ifndef __restrict__
#define __restrict__ __restrict
#endif
#define __launch_bounds__(...)
#define __shared__ static thread_local

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() = default; dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct uint3_ { unsigned x, y, z; };
extern thread_local uint3_ threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { memset(p, v, n); return 0; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n); return 0; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
#define CUDA_CHECK_THROW(x) do { (void)(x); } while (0)

inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float __powf(float a, float b) { return powf(a, b); }
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float rsqrt(float x) { return 1.0f / sqrtf(x); }  // CUDA host math: rsqrt(float) = 1 / sqrtf
inline unsigned __brev(unsigned x) {
	x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
	x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
	x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
	x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
	return (x >> 16) | (x << 16);
}
// the kernels run one "thread" at a time per OpenMP worker; atomics still have to be atomic across workers
template <typename T, typename U, typename = typename std::enable_if<std::is_integral<T>::value>::type> inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v) {
	uint32_t* ip = (uint32_t*)p; uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), neu; float f;
	do { memcpy(&f, &old, 4); f += v; memcpy(&neu, &f, 4); } while (!__atomic_compare_exchange_n(ip, &old, neu, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
	memcpy(&f, &old, 4); return f;
}
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED); while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return old; }

// kernel<<<blocks, threads, shmem, stream>>>(args...) is rewritten by oracle/ref_build.py into nsb_launch(blocks, threads, kernel, args...)
template <typename K, typename... Args> inline void nsb_launch(dim3 blocks, dim3 threads, K kernel, Args... args) {
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
	for (unsigned by = 0; by < blocks.y; ++by)
		for (unsigned bx = 0; bx < blocks.x; ++bx) {
			blockDim = threads; gridDim = blocks; blockIdx = {bx, by, 0};
			for (unsigned ty = 0; ty < threads.y; ++ty)
				for (unsigned tx = 0; tx < threads.x; ++tx) { threadIdx = {tx, ty, 0}; kernel(args...); }
		}
}

// CUDA's global min/max overload set (math_functions.hpp): mixed signedness resolves to unsigned, float/double kept
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
inline float min(float a, double b) { return fminf(a, (float)b); }
inline float max(float a, double b) { return fmaxf(a, (float)b); }
using std::signbit;
using std::abs;

// IEEE binary16 storage type with round-to-nearest-even conversions (the operations of cuda_fp16.h's __half used on this path)
struct __half {
	uint16_t bits = 0;
	__half() = default;
	__half(float f) { bits = from_float(f); }
	__half(double f) { bits = from_float((float)f); }
	__half(int i) { bits = from_float((float)i); }
	operator float() const { return to_float(bits); }
	__half& operator+=(const __half& o) { bits = from_float(to_float(bits) + to_float(o.bits)); return *this; }
	static uint16_t from_float(float f) {
		uint32_t x; memcpy(&x, &f, 4);
		uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7fffffffu;
		if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0));
		if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);  // rounds to inf
		if (x < 0x33000001u) return (uint16_t)sign;               // rounds to zero
		int e = (int)(x >> 23) - 127; uint32_t m = (x & 0x7fffffu) | 0x800000u;
		int shift; uint32_t base;
		if (e < -14) { shift = 13 + (-14 - e); base = 0; } else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
		uint32_t r = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
		if (rem > half || (rem == half && (r & 1u))) ++r;
		return (uint16_t)(sign | (base + r));
	}
	static float to_float(uint16_t h) {
		uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, x;
		if (e == 0) { if (m == 0) x = sign; else { int s = 0; while (!(m & 0x400u)) { m <<= 1; ++s; } x = sign | ((uint32_t)(113 - s) << 23) | ((m & 0x3ffu) << 13); } }
		else if (e == 31) x = sign | 0x7f800000u | (m << 13);
		else x = sign | ((e + 112) << 23) | (m << 13);
		float f; memcpy(&f, &x, 4); return f;
	}
};

#endif  // !__CUDACC__
