// tiny-cuda-nn/random.h + pcg32.h — TEST INFRASTRUCTURE ONLY (oracle/_ref): tcnn::default_rng_t is pcg32 (M. O'Neill's
// PCG-XSH-RR 64/32, public algorithm, Wenzel Jakob's pcg32.h API as vendored by tiny-cuda-nn).
#pragma once
#include <tiny-cuda-nn/common.h>
namespace tcnn {
struct pcg32 {
	static constexpr uint64_t DEFAULT_STATE = 0x853c49e6748fea9bULL, DEFAULT_STREAM = 0xda3e39cb94b95bdbULL, MULT = 0x5851f42d4c957f2dULL;
	uint64_t state, inc;
	TCNN_HOST_DEVICE pcg32() : state(DEFAULT_STATE), inc(DEFAULT_STREAM) {}
	TCNN_HOST_DEVICE pcg32(uint64_t initstate, uint64_t initseq = 1u) { seed(initstate, initseq); }
	TCNN_HOST_DEVICE void seed(uint64_t initstate, uint64_t initseq = 1) { state = 0U; inc = (initseq << 1u) | 1u; next_uint(); state += initstate; next_uint(); }
	TCNN_HOST_DEVICE uint32_t next_uint() {
		uint64_t oldstate = state;
		state = oldstate * MULT + inc;
		uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
		uint32_t rot = (uint32_t)(oldstate >> 59u);
		return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
	}
	TCNN_HOST_DEVICE float next_float() { union { uint32_t u; float f; } x; x.u = (next_uint() >> 9) | 0x3f800000u; return x.f - 1.0f; }
	TCNN_HOST_DEVICE void advance(int64_t delta_ = (1ll << 32)) {
		uint64_t cur_mult = MULT, cur_plus = inc, acc_mult = 1u, acc_plus = 0u, delta = (uint64_t)delta_;
		while (delta > 0) {
			if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
			cur_plus = (cur_mult + 1) * cur_plus; cur_mult *= cur_mult; delta /= 2;
		}
		state = acc_mult * state + acc_plus;
	}
};
using default_rng_t = pcg32;
}  // namespace tcnn
