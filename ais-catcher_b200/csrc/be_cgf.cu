// be_cgf.cu -- ModelDefault back end, first half: SquareFreqOffsetCorrection + FilterComplex (DSP.cpp:215-246, 417-489).
#include "exact.cuh"
#include "params.h"

namespace aisgpu {

// ---------------------------------------------------------------------------------------------
// K2a: SquareFreqOffsetCorrection, estimation half (DSP.cpp:417-455, FFT.h:93-130).
// One warp per 512-sample block: x^2 in bit-reversed order, the reference's radix-2 DIT butterflies stage by
// stage, |F| in fftshift order; then (one lane per block) the sequential float cumsum, then the parallel
// first-maximum searches.  Result: an index into the host-built phasor-step table.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CGF_THREADS) k_cgf_estimate(const float2 *__restrict__ Cbuf, long long c_stride, int c_begin, int nblk,
																 int total_blocks, const float2 *__restrict__ omega_g, int wide,
																 int *__restrict__ stepidx) {
	extern __shared__ __align__(16) unsigned char cgf_sm[];
	float2 *omega = reinterpret_cast<float2 *>(cgf_sm);                      // 512 float2
	float *mag = reinterpret_cast<float *>(cgf_sm + 4096);                   // [16][513]
	unsigned char *scratch = cgf_sm + 4096 + CGF_BLK_PER_CTA * CGF_ROWP * 4; // fft buffers, later cumsum [16][513]
	float2 *fftbuf = reinterpret_cast<float2 *>(scratch);
	float *cum = reinterpret_cast<float *>(scratch);

	const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
	for (int i = tid; i < CGF_N; i += CGF_THREADS) omega[i] = omega_g[i];
	__syncthreads();

	const int blk0 = blockIdx.x * CGF_BLK_PER_CTA;
	float2 *x = fftbuf + w * CGF_N;
	for (int rep = 0; rep < 2; rep++) {
		const int lb = w + rep * 8;
		const int id = blk0 + lb;
		if (id < total_blocks) {
			const int row = id / nblk, b = id - row * nblk;
			const float2 *src = Cbuf + (long long)row * c_stride + c_begin + (long long)b * CGF_N;
			for (int i = lane; i < CGF_N; i += 32) {
				float2 v = src[i];
				x[__brev((unsigned)i) >> 23] = cmul(v, v);
			}
			__syncwarp();
			for (int s = 0; s < 9; s++) {
				const int m2 = 1 << s;
				for (int q = lane; q < 256; q += 32) {
					const int j = q & (m2 - 1);
					const int lo = ((q >> s) << (s + 1)) + j, hi = lo + m2;
					const float2 o = omega[j << (8 - s)];
					const float2 t = cmul(o, x[hi]);
					const float2 a = x[lo];
					x[hi] = csub(a, t);
					x[lo] = cadd(a, t);
				}
				__syncwarp();
			}
			float *mg = mag + lb * CGF_ROWP;
			for (int i = lane; i < CGF_N; i += 32) mg[i] = habs(x[(i + 256) & 511]);
		}
		__syncwarp();
	}
	__syncthreads(); // all FFT buffers dead, mags complete
	if (wide && tid < CGF_BLK_PER_CTA && blk0 + tid < total_blocks) {
		const float *mg = mag + tid * CGF_ROWP;
		float *cs = cum + tid * CGF_ROWP;
		float c = 0.0f;
		cs[0] = 0.0f;
#pragma unroll 16
		for (int i = 1; i < CGF_N; i++) {
			c = __fadd_rn(c, mg[i]);
			cs[i] = c;
		}
	}
	__syncthreads();
	for (int rep = 0; rep < 2; rep++) {
		const int lb = w + rep * 8;
		const int id = blk0 + lb;
		if (id >= total_blocks) continue;
		const float *mg = mag + lb * CGF_ROWP;
		const float *cs = cum + lb * CGF_ROWP;
		int wi = 0;
		if (wide) { // DSP.cpp:424-446: M = 133, ofs = 15, delta = 102
			float bv = -1.0f;
			int bi = 1 << 30;
			for (int i = lane; i < CGF_N - 133; i += 32) {
				float v = __fadd_rn(__fsub_rn(cs[i + 133], cs[i]), __fmul_rn(0.6f, __fadd_rn(mg[i + 15], mg[i + 117])));
				if (v > bv) { bv = v; bi = i; }
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				float ov = __shfl_xor_sync(0xffffffffu, bv, o);
				int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
			}
			wi = (bi == (1 << 30)) ? 0 : bi;
			wi = wi + 66 - 256;
		}
		// DSP.cpp:448-455: i in [wi+187, wi+223)
		float bv = 0.0f;
		int bi = 1 << 30;
		for (int c = lane; c < 36; c += 32) {
			const int i = wi + 187 + c;
			float h = __fadd_rn(mg[i & 511], mg[(i + 102) & 511]);
			if (h > bv) { bv = h; bi = i; }
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			float ov = __shfl_xor_sync(0xffffffffu, bv, o);
			int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
		}
		if (lane == 0) stepidx[id] = (bi == (1 << 30)) ? CGF_IDX_NONE : bi + CGF_IDX_OFFSET;
	}
}

// ---------------------------------------------------------------------------------------------
// K2b: the CGF derotation phasor chain (DSP.cpp:457-465): rot *= rot_step per sample, rot /= |rot| per block.
// Strictly sequential per (stream, channel); one thread per row, all rows in flight at once.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_cgf_rot(const int *__restrict__ stepidx, const float2 *__restrict__ steptab, float2 *__restrict__ rot_state,
												  float2 *__restrict__ rots, long long r_stride, int nblk, int rows) {
	// lane = row.  The phasors of 32 consecutive steps are staged in shared memory and written out row by row, so that
	// every store instruction covers 256 contiguous bytes (a store per step and lane would touch 32 separate sectors
	// and make the store unit, not the multiply chain, the pace).
	__shared__ float2 tile[32][33];
	const int lane = threadIdx.x;
	const int row0 = blockIdx.x * 32;
	const int row = row0 + lane;
	const bool act = row < rows;
	float2 rot = act ? rot_state[row] : make_float2(1.0f, 0.0f);
	for (int b = 0; b < nblk; b++) {
		const float2 st = act ? steptab[stepidx[row * nblk + b]] : make_float2(1.0f, 0.0f);
		for (int i0 = 0; i0 < CGF_N; i0 += 32) {
#pragma unroll
			for (int i = 0; i < 32; i++) {
				rot = cmul(rot, st);
				tile[lane][i] = rot;
			}
			__syncwarp();
#pragma unroll 8
			for (int r = 0; r < 32; r++)
				if (row0 + r < rows) rots[(long long)(row0 + r) * r_stride + b * CGF_N + i0 + lane] = tile[r][lane];
			__syncwarp();
		}
		rot = cnormalize(rot);
	}
	if (act) rot_state[row] = rot;
}

// ---------------------------------------------------------------------------------------------
// K2c: output[i] *= rot (DSP.cpp:462) fused with FilterComplex 17 taps (DSP.cpp:215-246, Filters.h:35-41).
// ---------------------------------------------------------------------------------------------
__constant__ float c_taps_coherent[FIRC_T];

__global__ void __launch_bounds__(FIRC_TILE) k_cgf_derot_fir(const float2 *__restrict__ Cbuf, long long c_stride, int c_begin,
																const float2 *__restrict__ rots, long long r_stride, int nE,
																const float2 *__restrict__ hist_old, float2 *__restrict__ hist_new,
																float2 *__restrict__ Ebuf, long long e_stride, int e_off,
																float2 *__restrict__ tap_cgf, long long tap_stride) {
	__shared__ float2 der[FIRC_TILE + FIRC_T - 1];
	const int row = blockIdx.y, t0 = blockIdx.x * FIRC_TILE, tid = threadIdx.x;
	for (int i = tid; i < FIRC_TILE + FIRC_T - 1; i += FIRC_TILE) {
		const int n = t0 + i - (FIRC_T - 1);
		float2 v = make_float2(0.f, 0.f);
		if (n < 0) v = hist_old[row * (FIRC_T - 1) + (FIRC_T - 1) + n];
		else if (n < nE) {
			v = cmul(Cbuf[(long long)row * c_stride + c_begin + n], rots[(long long)row * r_stride + n]);
			if (tap_cgf && i >= FIRC_T - 1) tap_cgf[(long long)row * tap_stride + n] = v;
		}
		der[i] = v;
	}
	__syncthreads();
	const int n = t0 + tid;
	if (n < nE) {
		float2 x = make_float2(0.f, 0.f);
#pragma unroll
		for (int k = 0; k < FIRC_T; k++) {
			const float2 dd = der[tid + k];
			x.x = __fadd_rn(x.x, __fmul_rn(c_taps_coherent[k], dd.x));
			x.y = __fadd_rn(x.y, __fmul_rn(c_taps_coherent[k], dd.y));
		}
		Ebuf[(long long)row * e_stride + e_off + n] = x;
	}
	if (t0 + FIRC_TILE >= nE) { // the CTA holding the end of the row saves the next history
		for (int i = tid; i < FIRC_T - 1; i += FIRC_TILE) {
			const int nn = nE - (FIRC_T - 1) + i; // nE >= 512
			hist_new[row * (FIRC_T - 1) + i] = der[nn - t0 + (FIRC_T - 1)];
		}
	}
}

// ---- launch entry points ----
cudaError_t cgf_init(const float *taps17) {
	cudaError_t e = cudaMemcpyToSymbol(c_taps_coherent, taps17, FIRC_T * sizeof(float));
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(k_cgf_estimate, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 + 2 * CGF_BLK_PER_CTA * CGF_ROWP * 4);
}
cudaError_t launch_cgf_estimate(const float2 *Cbuf, long long c_stride, int c_begin, int nblk, int total_blocks, const float2 *omega, int wide, int *stepidx,
								cudaStream_t s) {
	const int ctas = (total_blocks + CGF_BLK_PER_CTA - 1) / CGF_BLK_PER_CTA;
	const size_t smem = 4096 + 2 * (size_t)CGF_BLK_PER_CTA * CGF_ROWP * 4;
	k_cgf_estimate<<<ctas, CGF_THREADS, smem, s>>>(Cbuf, c_stride, c_begin, nblk, total_blocks, omega, wide, stepidx);
	return cudaGetLastError();
}
cudaError_t launch_cgf_rot(const int *stepidx, const float2 *steptab, float2 *rot_state, float2 *rots, long long r_stride, int nblk, int rows, cudaStream_t s) {
	k_cgf_rot<<<(rows + 31) / 32, 32, 0, s>>>(stepidx, steptab, rot_state, rots, r_stride, nblk, rows);
	return cudaGetLastError();
}
cudaError_t launch_cgf_derot_fir(const float2 *Cbuf, long long c_stride, int c_begin, const float2 *rots, long long r_stride, int nE, const float2 *hist_old,
								 float2 *hist_new, float2 *Ebuf, long long e_stride, int e_off, float2 *tap_cgf, long long tap_stride, int rows, cudaStream_t s) {
	dim3 grid((nE + FIRC_TILE - 1) / FIRC_TILE, rows);
	k_cgf_derot_fir<<<grid, FIRC_TILE, 0, s>>>(Cbuf, c_stride, c_begin, rots, r_stride, nE, hist_old, hist_new, Ebuf, e_stride, e_off, tap_cgf, tap_stride);
	return cudaGetLastError();
}

} // namespace aisgpu
