// fft512.cuh -- the reference's 512-point radix-2 DIT FFT (DSP/FFT.h:93-130) of the squared block, one warp per block, data
// in registers; shared by SquareFreqOffsetCorrection (be_cgf.cu) and V2::FreqOffset::Estimate (be_v2.cu).  Every
// translation unit that includes this gets its own copy of the twiddle table (set through fft512_set_omega).
#pragma once
#include "exact.cuh"
#include "params.h"

namespace aisgpu {

static __constant__ float2 c_cgf_omega[CGF_N / 2]; // Omega[s] = polar(1, -2 pi s / N), host-computed (FFT.h:81-83)
constexpr int CGF_TB = CGF_N + CGF_N / 16;  // transpose tile: one pad slot per 16 values

__device__ __forceinline__ constexpr int cgf_rev4(int m) { return ((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3); }

// FFT of x^2 for one block; leaves |F[(i + 256) & 511]| in mg[i].  HYPOT: std::abs(complex) as glibc's hypotf (the CGF block,
// DSP.cpp:436) or sqrtf(re^2 + im^2) (V2::FreqOffset::Estimate, V2Engine.cpp:69-72)
template <bool HYPOT>
__device__ __forceinline__ void cgf_fft_block(const float2 *__restrict__ src, float2 *__restrict__ tb, float *__restrict__ mg, int lane, const float2 (&tw)[15]) {
	float2 x[16];
	const int rl = (int)(__brev((unsigned)lane) >> 27);
#pragma unroll
	for (int m = 0; m < 16; m++) { // position p = 16 lane + e holds sample rev9(p) = rev5(lane) + 32 rev4(e)
		const float2 v = src[rl + 32 * m];
		x[cgf_rev4(m)] = cmul(v, v);
	}
#pragma unroll
	for (int s = 0; s < 4; s++) {
		const int m2 = 1 << s;
#pragma unroll
		for (int q = 0; q < 8; q++) {
			const int j = q & (m2 - 1);
			const int lo = ((q >> s) << (s + 1)) + j, hi = lo + m2;
			const float2 t = cmul(c_cgf_omega[j << (8 - s)], x[hi]);
			const float2 a = x[lo];
			x[hi] = csub(a, t);
			x[lo] = cadd(a, t);
		}
	}
	{ // stage 4: positions p and p + 16 live in lanes l and l ^ 1, j = p & 15 = register index
		const bool odd = lane & 1;
#pragma unroll
		for (int e = 0; e < 16; e++) {
			const float2 mine = x[e];
			float2 other;
			other.x = __shfl_xor_sync(0xffffffffu, mine.x, 1);
			other.y = __shfl_xor_sync(0xffffffffu, mine.y, 1);
			const float2 hi = odd ? mine : other, lo = odd ? other : mine;
			float2 t = cmul(c_cgf_omega[e << 4], hi);
			if (odd) { t.x = -t.x; t.y = -t.y; } // x[hi] = a - t == a + (-t), exactly
			x[e] = cadd(lo, t);
		}
	}
	__syncwarp();
#pragma unroll
	for (int e = 0; e < 16; e++) tb[17 * lane + e] = x[e]; // slot p + (p >> 4), p = 16 lane + e
	__syncwarp();
#pragma unroll
	for (int r = 0; r < 16; r++) {
		const int p = lane + 32 * r;
		x[r] = tb[p + (p >> 4)];
	}
#pragma unroll
	for (int s = 5; s < 9; s++) { // positions p = lane + 32 r: bit s of p is bit s - 5 of r; j = p & (2^s - 1) = lane + 32 (r & (2^(s-5) - 1))
		const int sb = s - 5, m2r = 1 << sb;
#pragma unroll
		for (int q = 0; q < 8; q++) {
			const int jr = q & (m2r - 1);
			const int lo = ((q >> sb) << (sb + 1)) + jr, hi = lo + m2r;
			const float2 t = cmul(tw[m2r - 1 + jr], x[hi]);
			const float2 a = x[lo];
			x[hi] = csub(a, t);
			x[lo] = cadd(a, t);
		}
	}
#pragma unroll
	for (int r = 0; r < 16; r++)
		mg[(lane + 32 * r) ^ 256] = HYPOT ? habs(x[r]) : __fsqrt_rn(__fadd_rn(__fmul_rn(x[r].x, x[r].x), __fmul_rn(x[r].y, x[r].y)));
}


static inline cudaError_t fft512_set_omega(const float2 *omega256) { return cudaMemcpyToSymbol(c_cgf_omega, omega256, (CGF_N / 2) * sizeof(float2)); }
// per-lane twiddles of stages 5..8: Omega[(lane + 32 jr) << (3 - sb)] for stage 5 + sb, jr < 2^sb
__device__ __forceinline__ void fft512_lane_twiddles(const float2 *__restrict__ omega_g, int lane, float2 (&tw)[15]) {
#pragma unroll
	for (int sb = 0; sb < 4; sb++)
#pragma unroll
		for (int jr = 0; jr < (1 << sb); jr++) tw[(1 << sb) - 1 + jr] = omega_g[(lane + 32 * jr) << (3 - sb)];
}

} // namespace aisgpu
