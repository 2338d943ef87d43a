// aisgpu_kernels.cuh -- sm_100a kernels of the AIS demodulation hot path.
//
// Every kernel reproduces the IEEE binary32 operation order of the reference block it replaces
// (file:line cited per kernel, relative to /root/reference/Source) so that results are bit-identical:
// all arithmetic goes through __fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn (never contracted to FMA), std::abs of a
// complex is evaluated as glibc's hypotf does ((float)sqrt((double)x*x+(double)y*y)), libm-dependent
// constants (twiddles, phasor steps) come from host tables, and atan2f is the fdlibm algorithm glibc 2.39 ships.
//
// Layout in HBM (one engine == one batch of B independent IQ streams, "row" = stream*2 + channel):
//   in     [B][N]              input samples of one submit (CF32 float2, or CU8/CS8/CS16)
//   tail   [B][P]              last P input samples of the previous submit (front-end warm-up history)
//   rot    [P96 + N>>k]        Rotate phasor table of the submit (shared by all streams), with P96 history
//   Cbuf   [2B][HC + n48max]   48 kHz channel samples; new samples land at offset HC, unconsumed/history before
//   rots   [2B][nE]            CGF derotation phasors of the submit
//   Ebuf   [2B][HE + nEmax]    samples entering the symbol-timing stage (FIR17 out, or FIR37 out for FM models)
//   state  PS/decoder/CGF/FIR  small per-row / per-(row,phase) structs
//   frames ring of FrameRec    decoded frames of the submit
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace aisgpu {

// ---------------------------------------------------------------------------------------------
// exact arithmetic helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(__fmul_rn(a.x, s), __fmul_rn(a.y, s)); }
// std::complex<float> product (ac-bd, ad+bc), every product and sum rounded separately
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
	return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// std::abs(std::complex<float>) == cabsf == glibc 2.39 hypotf for finite inputs (checked on 5e7 patterns)
__device__ __forceinline__ float habs(float2 a) {
	double x = (double)a.x, y = (double)a.y;
	return __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y))));
}
__device__ __forceinline__ float2 cnormalize(float2 r) { // rot /= std::abs(rot)
	float a = habs(r);
	return make_float2(__fdiv_rn(r.x, a), __fdiv_rn(r.y, a));
}

// fdlibm atanf/atan2f (the algorithm behind glibc 2.39 __ieee754_atan2f; verified bit-identical on 5e7 inputs)
__device__ __forceinline__ float fd_atanf(float x) {
	const float aT[11] = { 3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
						   -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f };
	const int hx = __float_as_int(x), ix = hx & 0x7fffffff;
	if (ix >= 0x4c000000) { // |x| >= 2^25 (rare)
		if (ix > 0x7f800000) return __fadd_rn(x, x);
		const float r = __fadd_rn(1.5707962513e+00f, 7.5497894159e-08f);
		return hx > 0 ? r : -r;
	}
	if (ix < 0x31000000) return x; // |x| < 2^-29 (rare)
	// The four argument reductions of fdlibm are evaluated side by side and selected, so a warp does not serialise
	// over them; the operations on the selected path are exactly the library's.
	const bool small = ix < 0x3ee00000; // |x| < 0.4375: no reduction, x keeps its sign
	const float ax = fabsf(x);
	const int id = ix < 0x3f300000 ? 0 : (ix < 0x3f980000 ? 1 : (ix < 0x401c0000 ? 2 : 3));
	const float n0 = __fsub_rn(__fmul_rn(2.0f, ax), 1.0f), d0 = __fadd_rn(2.0f, ax);
	const float n1 = __fsub_rn(ax, 1.0f), d1 = __fadd_rn(ax, 1.0f);
	const float n2 = __fsub_rn(ax, 1.5f), d2 = __fadd_rn(1.0f, __fmul_rn(1.5f, ax));
	const float num = id == 0 ? n0 : (id == 1 ? n1 : (id == 2 ? n2 : -1.0f));
	const float den = id == 0 ? d0 : (id == 1 ? d1 : (id == 2 ? d2 : ax));
	const float hi = id == 0 ? 4.6364760399e-01f : (id == 1 ? 7.8539812565e-01f : (id == 2 ? 9.8279368877e-01f : 1.5707962513e+00f));
	const float lo = id == 0 ? 5.0121582440e-09f : (id == 1 ? 3.7748947079e-08f : (id == 2 ? 3.4473217170e-08f : 7.5497894159e-08f));
	const float xr = small ? x : __fdiv_rn(num, den);
	const float z = __fmul_rn(xr, xr), w = __fmul_rn(z, z);
	const float s1 = __fmul_rn(z, __fadd_rn(aT[0], __fmul_rn(w, __fadd_rn(aT[2], __fmul_rn(w, __fadd_rn(aT[4], __fmul_rn(w, __fadd_rn(aT[6], __fmul_rn(w, __fadd_rn(aT[8], __fmul_rn(w, aT[10])))))))))));
	const float s2 = __fmul_rn(w, __fadd_rn(aT[1], __fmul_rn(w, __fadd_rn(aT[3], __fmul_rn(w, __fadd_rn(aT[5], __fmul_rn(w, __fadd_rn(aT[7], __fmul_rn(w, aT[9])))))))));
	const float xs = __fmul_rn(xr, __fadd_rn(s1, s2));
	if (small) return __fsub_rn(xr, xs);
	const float r = __fsub_rn(hi, __fsub_rn(__fsub_rn(xs, lo), xr));
	return hx < 0 ? -r : r;
}
__device__ __forceinline__ float fd_atan2f(float y, float x) {
	const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
	int hx = __float_as_int(x), ix = hx & 0x7fffffff, hy = __float_as_int(y), iy = hy & 0x7fffffff;
	if (ix > 0x7f800000 || iy > 0x7f800000) return __fadd_rn(x, y);
	if (hx == 0x3f800000) return fd_atanf(y);
	int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
	if (iy == 0) {
		switch (m) {
		case 0: case 1: return y;
		case 2: return __fadd_rn(pi, tiny);
		default: return __fsub_rn(-pi, tiny);
		}
	}
	if (ix == 0) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
	if (ix == 0x7f800000) {
		if (iy == 0x7f800000) {
			switch (m) {
			case 0: return __fadd_rn(pi_o_4, tiny);
			case 1: return __fsub_rn(-pi_o_4, tiny);
			case 2: return __fadd_rn(__fmul_rn(3.0f, pi_o_4), tiny);
			default: return __fsub_rn(__fmul_rn(-3.0f, pi_o_4), tiny);
			}
		}
		else {
			switch (m) {
			case 0: return 0.0f;
			case 1: return -0.0f;
			case 2: return __fadd_rn(pi, tiny);
			default: return __fsub_rn(-pi, tiny);
			}
		}
	}
	if (iy == 0x7f800000) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
	int k = (iy - ix) >> 23;
	float z;
	if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
	else if (hx < 0 && k < -60) z = 0.0f;
	else z = fd_atanf(fabsf(__fdiv_rn(y, x)));
	switch (m) {
	case 0: return z;
	case 1: return __int_as_float(__float_as_int(z) ^ 0x80000000);
	case 2: return __fsub_rn(pi, __fsub_rn(z, pi_lo));
	default: return __fsub_rn(__fsub_rn(z, pi_lo), pi);
	}
}

// ---------------------------------------------------------------------------------------------
// K0: Rotate phasor table (DSP/DSP.cpp:296-316: rot *= mult per 96 kHz sample, rot /= |rot| once per call)
// The phasor depends only on the sequence of chunk lengths, never on the data, so one table per submit serves
// every stream of the batch.  tab[P96 + i] is the phasor that multiplies 96 kHz sample i of this submit;
// tab[0..P96) repeats the last P96 phasors of the previous submit (warm-up history of the front end).
// ---------------------------------------------------------------------------------------------
__global__ void k_rot_table(float2 *__restrict__ tab, const float2 *__restrict__ prev_tail, const float2 *__restrict__ state_in,
							float2 *__restrict__ state_out, float2 mult, int P96, int n96) {
	if (blockIdx.x != 0) return;
	for (int i = threadIdx.x; i < P96; i += blockDim.x) tab[i] = prev_tail ? prev_tail[i] : make_float2(1.0f, 0.0f);
	if (threadIdx.x != 0) return;
	float2 rot = *state_in;
	float2 *o = tab + P96;
	for (int i = 0; i < n96; i++) {
		o[i] = rot;
		rot = cmul(rot, mult);
	}
	*state_out = cnormalize(rot);
}

// ---------------------------------------------------------------------------------------------
// K1: fused front end.  input rate -> k x Downsample2CIC5 (DSP.cpp:93-117) -> FilterComplex3Tap (DSP.cpp:283-293)
//     -> Rotate (DSP.cpp:296-316) -> per channel Downsample2CIC5 -> FilterCIC5 (DSP.cpp:132-157) -> Cbuf.
// One CTA owns (segment, stream) and walks the segment tile by tile, every stage array living in shared memory
// as [HIST history | tile]; the history is what the reference keeps in h0..h4 / h1,h2 / rot.  A segment starts P
// samples early (from the previous submit's tail for segment 0) with zero history: after P >= h_k samples every
// stage's history is exact because each CIC stage is a pure function of its last 6 inputs
// (u_{s+1}[n] = fl(u_s[n] + u_s[n-1]), y[j] = u_5[2j]/32).
// ---------------------------------------------------------------------------------------------
constexpr int FE_HIST = 6;   // >= 5, even so that even sample indices stay 16-byte aligned
constexpr int FE_SLACK = 12; // over-read room behind each array for partial runs
constexpr int FE_MAXK = 7;

struct FeParams {
	const void *in;       // [B][in_stride] samples
	const void *tail;     // [B][P]
	long long in_stride;  // in samples
	int format, k, N, P, seg_len, tile;
	int use_fdc;
	float fdc_alpha, fdc_beta;
	const float2 *rot;    // [P96 + N>>k]
	float2 *C;            // [2B][c_stride]
	long long c_stride;
	int c_off;
	int off_in[2];           // smem offsets (float2 units) of the two input-ring buffers (level 0, each [HIST | tile])
	int off_rot[2];          // phasors of the tile
	int off_lv[FE_MAXK + 1]; // level arrays 1..k (off_lv[0] unused)
	int off_up, off_dn, off_wa, off_wb;
	int off_l1b, off_l3b; // warp-specialised kernel: second buffers of the level-1 and level-3 arrays
	int st_S, st_wps, st_B; // streaming kernel: samples per lane sub-segment, warps per stream, streams
	int smem_f2;          // total float2
	float2 *D0;           // PRE mode (decimation in front of DSP::Upsample): level-K samples, [B][d0_stride], sample i at d0_off + i
	long long d0_stride;
	int d0_off;
};

template <int FMT>
__device__ __forceinline__ void fe_load_pair(const void *base, long long idx, float2 &a, float2 &b) {
	// two consecutive samples starting at even index idx
	if (FMT == 0) {
		float4 v = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(base) + idx));
		a = make_float2(v.x, v.y);
		b = make_float2(v.z, v.w);
	}
	else if (FMT == 1) { // CU8: (u-128)/128  (Utilities/Convert.cpp:255-264); /128 is an exact scaling
		uchar4 v = __ldg(reinterpret_cast<const uchar4 *>(reinterpret_cast<const uchar2 *>(base) + idx));
		a = make_float2(__fmul_rn((float)((int)v.x - 128), 0.0078125f), __fmul_rn((float)((int)v.y - 128), 0.0078125f));
		b = make_float2(__fmul_rn((float)((int)v.z - 128), 0.0078125f), __fmul_rn((float)((int)v.w - 128), 0.0078125f));
	}
	else if (FMT == 2) { // CS8 (Convert.cpp:266-275)
		char4 v = __ldg(reinterpret_cast<const char4 *>(reinterpret_cast<const char2 *>(base) + idx));
		a = make_float2(__fmul_rn((float)v.x, 0.0078125f), __fmul_rn((float)v.y, 0.0078125f));
		b = make_float2(__fmul_rn((float)v.z, 0.0078125f), __fmul_rn((float)v.w, 0.0078125f));
	}
	else { // CS16 (Convert.cpp:277-286)
		short4 v = __ldg(reinterpret_cast<const short4 *>(reinterpret_cast<const short2 *>(base) + idx));
		a = make_float2(__fmul_rn((float)v.x, 3.0517578125e-05f), __fmul_rn((float)v.y, 3.0517578125e-05f));
		b = make_float2(__fmul_rn((float)v.z, 3.0517578125e-05f), __fmul_rn((float)v.w, 3.0517578125e-05f));
	}
}

// ---- packed binary32 pairs (SASS FADD2 / FMUL2): one instruction rounds both lanes exactly like two scalar
// ---- __fadd_rn / __fmul_rn (verified bit-for-bit on 1.6e7 patterns incl. denormals, tools/microbench_f32x2.cu);
// ---- a complex sample is one 64-bit register pair, so a complex add is ONE issue slot instead of two.
typedef unsigned long long c64;
__device__ __forceinline__ c64 padd(c64 a, c64 b) {
	c64 r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ c64 pmul(c64 a, c64 b) {
	c64 r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ c64 pack2(float x, float y) {
	c64 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
	return r;
}
__device__ __forceinline__ float2 unpack2(c64 v) {
	float2 r;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
	return r;
}

// R consecutive outputs of one Downsample2CIC5 from 2R+4 inputs held in registers: 9R+6 complex adds.
// sm = the CTA's shared array; in_off / out_off = index of sample 0 of the stage input / output (history at
// negative indices); j0 = first output index.  Outputs past the valid count land in the arrays' slack.
template <int R>
__device__ __forceinline__ void ds2_run(float2 *__restrict__ sm, int in_off, int out_off, int j0) {
	c64 v[2 * R + 6];
	const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(sm + in_off + 2 * j0 - 6);
#pragma unroll
	for (int q = R + 2; q >= 0; q--) {
		const ulonglong2 t = p[q];
		v[2 * q] = t.x;
		v[2 * q + 1] = t.y;
	}
#pragma unroll
	for (int s = 1; s <= 4; s++) {
#pragma unroll
		for (int n = 2 * R + 4; n >= s + 1; n--) v[n] = padd(v[n], v[n - 1]);
	}
	const c64 sc = pack2(0.03125f, 0.03125f);
	c64 *o = reinterpret_cast<c64 *>(sm + out_off + j0);
#pragma unroll
	for (int q = 0; q < R; q++) {
		const int n = 6 + 2 * q;
		o[q] = pmul(padd(v[n], v[n - 1]), sc);
	}
}

// R consecutive outputs of FilterCIC5 (no decimation) from R+5 inputs: 5R+10 complex adds; straight to HBM.
template <int R>
__device__ __forceinline__ void fcic_run(const float2 *__restrict__ sm, int in_off, float2 *__restrict__ out, int m0, int m_lo, int n_out) {
	c64 v[R + 5];
	const c64 *p = reinterpret_cast<const c64 *>(sm + in_off + m0 - 5);
#pragma unroll
	for (int q = 0; q < R + 5; q++) v[q] = p[q];
#pragma unroll
	for (int s = 1; s <= 5; s++) {
#pragma unroll
		for (int n = R + 4; n >= s; n--) v[n] = padd(v[n], v[n - 1]);
	}
	const c64 sc = pack2(0.03125f, 0.03125f);
	c64 *o = reinterpret_cast<c64 *>(out);
#pragma unroll
	for (int q = 0; q < R; q++)
		if (m0 + q >= m_lo && m0 + q < n_out) o[m0 + q] = pmul(v[5 + q], sc);
}

// ---- mbarrier + 1-D bulk async copy (TMA, SASS UBLKCP): global -> shared without touching registers ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
		"r"(parity)
		: "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
					 (unsigned)__cvta_generic_to_shared(smem_dst)),
				 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
				 : "memory");
}

// One CTA of NW warps owns (segment, stream) and walks [seg_start - P, seg_end) in tiles of p.tile input samples,
// starting from zero history: after P >= h_k samples every stage's history is exact, so only 48 kHz outputs that
// belong to [seg_start, seg_end) are written.  Thread 0 keeps a two-deep ring of bulk async copies (input tile + its
// Rotate phasors) in flight; all threads then run the stages of the tile back to back out of the CTA's shared-memory
// arrays ([HIST history | tile] each).  The arrays of one CTA serve NW warps, so the shared-memory footprint per
// resident warp -- what capped the one-warp version at 6 warps per SM -- drops NW-fold; the deeper (shorter) stages
// simply occupy fewer warps.  After the barrier that ends a stage, three threads move the last HIST inputs of that
// stage to the front of the array the next tile will read (the reference's h0..h4 / h1,h2 carried state).
template <int NW>
__device__ __forceinline__ void fe_sync() {
	if (NW == 1) __syncwarp();
	else __syncthreads();
}

// History for the next tile: dst[-HIST .. 0) = src[n - HIST .. n) for a group of stage arrays, n = len >> shift (even;
// when n < HIST part of the old history moves up -- one warp instruction loads all entries before any is stored).
// Array descriptors (src offset, dst offset, shift) sit in shared memory; three threads of warp 0 serve one array.
struct FeCarryDesc { int src, dst, shift, pad; };
__device__ __forceinline__ void fe_carry_group(float2 *__restrict__ sm, const FeCarryDesc *__restrict__ desc, int first, int count, int len, int tid) {
	if (tid < 3 * count) {
		const int a = tid / 3, e = 2 * (tid - 3 * a);
		const FeCarryDesc d = desc[first + a];
		const int n = len >> d.shift;
		const float4 v = *reinterpret_cast<const float4 *>(sm + d.src + n - FE_HIST + e);
		*reinterpret_cast<float4 *>(sm + d.dst - FE_HIST + e) = v;
	}
}

template <int FMT, int NW, int K, bool PRE = false>
__global__ void __launch_bounds__(NW * 32) k_frontend(const FeParams p) {
	constexpr int NT = NW * 32;
	extern __shared__ __align__(16) float2 sm[];
	__shared__ __align__(8) uint64_t mbar[2];
	__shared__ FeCarryDesc cdesc[2][FE_MAXK + 6]; // [parity of the tile][array]: input ring, levels 1..K, up, dn | wa, wb
	const int tid = threadIdx.x;
	const int stream = blockIdx.y;
	const long long seg_start = (long long)blockIdx.x * p.seg_len;
	if (seg_start >= p.N) return;
	const int seg_n = (int)min((long long)p.seg_len, (long long)p.N - seg_start); // samples of this segment
	const int span = seg_n + p.P;                                                  // samples walked, warm-up included
	const int n_tiles = (span + p.tile - 1) / p.tile;
	const long long base = seg_start - p.P; // first sample walked, relative to the submit (negative: previous submit's tail)
	const float2 *rot_g = p.rot + (p.P >> K) + (base >> K);
	float2 *Cg = p.C + (long long)(stream * 2) * p.c_stride + p.c_off + (base >> (K + 1));
	const int m_first = p.P >> (K + 1); // first 48 kHz output (relative to base) that belongs to the segment
	const int off_up = p.off_up + FE_HIST, off_dn = p.off_dn + FE_HIST, off_wa = p.off_wa + FE_HIST, off_wb = p.off_wb + FE_HIST;

	// zero what acts as history or may be read before written (the whole array is small enough to clear)
	for (int i = tid; i < p.smem_f2; i += NT) sm[i] = make_float2(0.f, 0.f);
	if (tid == 0) {
		mbar_init(&mbar[0], 1);
		mbar_init(&mbar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (tid < 2 * (K + 5)) {
		const int par = tid / (K + 5), a = tid - par * (K + 5);
		FeCarryDesc d;
		d.pad = 0;
		if (a == 0) { d.src = p.off_in[par] + FE_HIST; d.dst = p.off_in[par ^ 1] + FE_HIST; d.shift = 0; }
		else if (a <= K) { d.src = d.dst = p.off_lv[a] + FE_HIST; d.shift = a; }
		else if (a == K + 1) { d.src = d.dst = off_up; d.shift = K; }
		else if (a == K + 2) { d.src = d.dst = off_dn; d.shift = K; }
		else if (a == K + 3) { d.src = d.dst = off_wa; d.shift = K + 1; }
		else { d.src = d.dst = off_wb; d.shift = K + 1; }
		cdesc[par][a] = d;
	}
	fe_sync<NW>();

	auto issue = [&](int t) {
		const int rel = t * p.tile;
		const int len = min(p.tile, span - rel);
		const int b = t & 1;
		const int n96 = len >> K;
		unsigned bytes = PRE ? 0u : (unsigned)n96 * 8u;
		if (FMT == 0) bytes += (unsigned)len * 8u;
		mbar_expect_tx(&mbar[b], bytes);
		if (!PRE) bulk_g2s(sm + p.off_rot[b], rot_g + (rel >> K), (unsigned)n96 * 8u, &mbar[b]);
		if (FMT == 0) {
			const long long pos = base + rel;
			float2 *dst = sm + p.off_in[b] + FE_HIST;
			const float2 *in = reinterpret_cast<const float2 *>(p.in) + (long long)stream * p.in_stride;
			const float2 *tl = reinterpret_cast<const float2 *>(p.tail) + (long long)stream * p.P + p.P;
			if (pos >= 0) bulk_g2s(dst, in + pos, (unsigned)len * 8u, &mbar[b]);
			else if (pos + len <= 0) bulk_g2s(dst, tl + pos, (unsigned)len * 8u, &mbar[b]);
			else { // the tile straddles the first sample of the submit
				const int nt = (int)(-pos);
				bulk_g2s(dst, tl + pos, (unsigned)nt * 8u, &mbar[b]);
				bulk_g2s(dst + nt, in, (unsigned)(len - nt) * 8u, &mbar[b]);
			}
		}
	};
	if (tid == 0) issue(0);

	for (int t = 0; t < n_tiles; t++) {
		const int rel = t * p.tile;
		const int len = min(p.tile, span - rel);
		const int b = t & 1;
		if (tid == 0 && t + 1 < n_tiles) issue(t + 1); // ring slot b^1 was released by the barriers of tile t-1
		const int off_in = (b ? p.off_in[1] : p.off_in[0]) + FE_HIST;
		if (FMT != 0) { // integer formats: convert while loading (registers), no bulk copy
			const long long pos = base + rel;
			const long long tbase = (long long)stream * p.P + p.P + pos;
			const long long ibase = (long long)stream * p.in_stride + pos;
			for (int i = tid * 2; i < len; i += 2 * NT) {
				float2 x, y;
				if (pos + i < 0) fe_load_pair<FMT>(p.tail, tbase + i, x, y);
				else fe_load_pair<FMT>(p.in, ibase + i, x, y);
				*reinterpret_cast<float4 *>(sm + off_in + i) = make_float4(x.x, x.y, y.x, y.y);
			}
			fe_sync<NW>();
		}
		mbar_wait(&mbar[b], (unsigned)((t >> 1) & 1));
		// The deeper stages only have work for one or two warps.  Warp w of every CTA sits on scheduler w % 4, so a fixed
		// assignment would pile all of that work on one of the SM's four schedulers; the work index vt is therefore
		// rotated by one warp per stage and per tile, which spreads it evenly (CTAs are at different tiles).
		int rotw = t + blockIdx.x;
#define FE_VT() ((tid + 32 * ((rotw++) & (NW - 1))) & (NT - 1))
		// ---- K cascaded Downsample2CIC5 at the input rate ----
		int src = off_in;
#pragma unroll
		for (int l = 0; l < K; l++) {
			const int dst = p.off_lv[l + 1] + FE_HIST;
			const int n_out = len >> (l + 1);
			const int vt = FE_VT();
			for (int j0 = vt * 5; j0 < n_out; j0 += 5 * NT) ds2_run<5>(sm, src, dst, j0);
			fe_sync<NW>();
			if (l == 0 && !PRE) fe_carry_group(sm, cdesc[b], K + 3, 2, p.tile, tid); // wa, wb of the previous (always full) tile; its FilterCIC5 pass is two barriers back
			src = dst;
		}
		if (PRE) { // decimation in front of DSP::Upsample (Model.cpp:183-189): the level-K samples go to HBM
			const int nK = len >> K, iK = rel >> K, firstK = p.P >> K; // samples before firstK are warm-up
			float2 *o = p.D0 + (long long)stream * p.d0_stride + p.d0_off + (base >> K) + iK;
			const int vt_o = FE_VT();
			for (int i = vt_o; i < nK; i += NT)
				if (iK + i >= firstK) o[i] = sm[src + i];
			fe_sync<NW>();
			fe_carry_group(sm, cdesc[b], 0, K + 1, len, tid);
			fe_sync<NW>();
			continue;
		}
		// ---- FilterComplex3Tap + Rotate at 96 kHz ----
		const int n96 = len >> K;
		const int off_rt = b ? p.off_rot[1] : p.off_rot[0];
		const int vt_r = FE_VT();
		for (int i = vt_r; i < n96; i += NT) {
			float2 x = sm[src + i];
			if (p.use_fdc) { // alpha * (h1 + data[i]) + h2 * beta
				const float2 tt = cadd(sm[src + i - 2], x);
				const float2 h2 = sm[src + i - 1];
				x = make_float2(__fadd_rn(__fmul_rn(p.fdc_alpha, tt.x), __fmul_rn(h2.x, p.fdc_beta)),
								__fadd_rn(__fmul_rn(p.fdc_alpha, tt.y), __fmul_rn(h2.y, p.fdc_beta)));
			}
			const float2 r = sm[off_rt + i];
			const float RR = __fmul_rn(x.x, r.x), II = __fmul_rn(x.y, r.y), RI = __fmul_rn(x.x, r.y), IR = __fmul_rn(x.y, r.x);
			sm[off_up + i] = make_float2(__fsub_rn(RR, II), __fadd_rn(IR, RI));
			sm[off_dn + i] = make_float2(__fadd_rn(RR, II), __fsub_rn(IR, RI));
		}
		fe_sync<NW>();
		if (K == 0) fe_carry_group(sm, cdesc[b], K + 3, 2, p.tile, tid);
		// ---- per channel Downsample2CIC5 96k -> 48k ----
		const int n48 = n96 >> 1;
		const int runs = (n48 + 4) / 5;
		if (K == 0) fe_sync<NW>(); // the wa/wb history move above reads what this pass overwrites
		const int vt_c = FE_VT();
		for (int r = vt_c; r < 2 * runs; r += NT) {
			const int ch = r >= runs;
			ds2_run<5>(sm, ch ? off_dn : off_up, ch ? off_wb : off_wa, (ch ? r - runs : r) * 5);
		}
		fe_sync<NW>();
		// every stage that reads the input ring, the level arrays, up and dn has run: move their histories
		fe_carry_group(sm, cdesc[b], 0, K + 3, len, tid);
		// ---- per channel FilterCIC5 at 48k, straight to HBM ----
		const int m_rel = rel >> (K + 1); // 48 kHz index of the tile's first output, relative to base
		if (m_rel + n48 > m_first) {
			const int m_lo = m_first - m_rel; // outputs before it are warm-up
			const int vt_f = FE_VT();
			for (int r = vt_f; r < 2 * runs; r += NT) {
				const int ch = r >= runs;
				fcic_run<5>(sm, ch ? off_wb : off_wa, Cg + (ch ? p.c_stride : 0) + m_rel, (ch ? r - runs : r) * 5, m_lo, n48);
			}
		}
		fe_sync<NW>();
	}
}
#undef FE_VT

// ---------------------------------------------------------------------------------------------
// K1': the same front end as a three-stage software pipeline of specialised warps (K >= 3, i.e. >= 768 kS/s).
// In k_frontend the deeper stages keep one warp busy while the other three wait at the CTA barrier, so an SM rarely
// has more than one runnable warp per resident CTA.  Here every warp owns a group of stages and the groups work on
// different tiles at the same time:
//   warp 0: input tile (TMA ring)            -> Downsample2CIC5 level 1                -> L1[t & 1]
//   warp 1: L1[t & 1]                        -> levels 2, 3                             -> L3[t & 1]
//   warp 2: L3[t & 1] -> levels 4..K -> FilterComplex3Tap -> Rotate -> 2 x (Downsample2CIC5, FilterCIC5) -> HBM
// hand-offs through mbarriers (full/empty per double buffer); inside a warp the stages are ordered by __syncwarp.
// The arithmetic (ds2_run / fcic_run, operation order) is exactly that of k_frontend.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// history for the next tile, moved by three lanes of the calling warp (see fe_carry_group); n even
__device__ __forceinline__ void ws_carry(float2 *__restrict__ sm, int src, int dst, int n, int lane) {
	if (lane < FE_HIST / 2) {
		const float4 v = *reinterpret_cast<const float4 *>(sm + src + n - FE_HIST + 2 * lane);
		*reinterpret_cast<float4 *>(sm + dst - FE_HIST + 2 * lane) = v;
	}
}
constexpr int WS_THREADS = 96;
template <int FMT, int K>
__global__ void __launch_bounds__(WS_THREADS) k_frontend_ws(const FeParams p) {
	static_assert(K >= 3, "the warp-specialised front end needs at least three CIC stages");
	extern __shared__ __align__(16) float2 sm[];
	__shared__ __align__(8) uint64_t bar_in[2], bar_f1[2], bar_e1[2], bar_f3[2], bar_e3[2];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int stream = blockIdx.y;
	const long long seg_start = (long long)blockIdx.x * p.seg_len;
	if (seg_start >= p.N) return;
	const int seg_n = (int)min((long long)p.seg_len, (long long)p.N - seg_start);
	const int span = seg_n + p.P;
	const int n_tiles = (span + p.tile - 1) / p.tile;
	const long long base = seg_start - p.P;
	for (int i = tid; i < p.smem_f2; i += WS_THREADS) sm[i] = make_float2(0.f, 0.f);
	if (tid == 0) {
		for (int b = 0; b < 2; b++) {
			mbar_init(&bar_in[b], 1);
			mbar_init(&bar_f1[b], 1);
			mbar_init(&bar_e1[b], 1);
			mbar_init(&bar_f3[b], 1);
			mbar_init(&bar_e3[b], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const int L1_0 = p.off_lv[1] + FE_HIST, L1_1 = p.off_l1b + FE_HIST, L3_0 = p.off_lv[3] + FE_HIST, L3_1 = p.off_l3b + FE_HIST;
#define L1(bb) ((bb) ? L1_1 : L1_0)
#define L3(bb) ((bb) ? L3_1 : L3_0)

	if (warp == 0) {
		// ---------------- input ring -> level 1 ----------------
		auto issue = [&](int t) {
			const int rel = t * p.tile;
			const int len = min(p.tile, span - rel);
			const int b = t & 1;
			mbar_expect_tx(&bar_in[b], (unsigned)len * 8u);
			const long long pos = base + rel;
			float2 *dst = sm + p.off_in[b] + FE_HIST;
			const float2 *in = reinterpret_cast<const float2 *>(p.in) + (long long)stream * p.in_stride;
			const float2 *tl = reinterpret_cast<const float2 *>(p.tail) + (long long)stream * p.P + p.P;
			if (pos >= 0) bulk_g2s(dst, in + pos, (unsigned)len * 8u, &bar_in[b]);
			else if (pos + len <= 0) bulk_g2s(dst, tl + pos, (unsigned)len * 8u, &bar_in[b]);
			else {
				const int nt = (int)(-pos);
				bulk_g2s(dst, tl + pos, (unsigned)nt * 8u, &bar_in[b]);
				bulk_g2s(dst + nt, in, (unsigned)(len - nt) * 8u, &bar_in[b]);
			}
		};
		if (FMT == 0 && lane == 0) issue(0);
		for (int t = 0; t < n_tiles; t++) {
			const int rel = t * p.tile;
			const int len = min(p.tile, span - rel);
			const int b = t & 1;
			const int off_in = p.off_in[b] + FE_HIST, off_in_next = p.off_in[b ^ 1] + FE_HIST;
			if (FMT == 0) {
				if (lane == 0 && t + 1 < n_tiles) issue(t + 1); // slot b^1 was released when this warp finished tile t-1
				mbar_wait(&bar_in[b], (unsigned)((t >> 1) & 1));
			}
			else {
				const long long pos = base + rel;
				const long long tbase = (long long)stream * p.P + p.P + pos;
				const long long ibase = (long long)stream * p.in_stride + pos;
				for (int i = lane * 2; i < len; i += 64) {
					float2 x, y;
					if (pos + i < 0) fe_load_pair<FMT>(p.tail, tbase + i, x, y);
					else fe_load_pair<FMT>(p.in, ibase + i, x, y);
					*reinterpret_cast<float4 *>(sm + off_in + i) = make_float4(x.x, x.y, y.x, y.y);
				}
				__syncwarp();
			}
			if (t >= 2) mbar_wait(&bar_e1[b], (unsigned)(((t >> 1) + 1) & 1)); // warp 1 is done with L1(b) of tile t-2
			const int n_out = len >> 1;
			for (int j0 = lane * 5; j0 < n_out; j0 += 160) ds2_run<5>(sm, off_in, L1(b), j0);
			__syncwarp();
			ws_carry(sm, off_in, off_in_next, len, lane);
			__syncwarp();
			if (lane == 0) mbar_arrive(&bar_f1[b]);
		}
	}
	else if (warp == 1) {
		// ---------------- level 1 -> levels 2, 3 ----------------
		const int L2 = p.off_lv[2] + FE_HIST;
		for (int t = 0; t < n_tiles; t++) {
			const int len = min(p.tile, span - t * p.tile);
			const int b = t & 1;
			mbar_wait(&bar_f1[b], (unsigned)((t >> 1) & 1));
			const int n2 = len >> 2;
			for (int j0 = lane * 5; j0 < n2; j0 += 160) ds2_run<5>(sm, L1(b), L2, j0);
			__syncwarp();
			ws_carry(sm, L1(b), L1(b ^ 1), len >> 1, lane); // history of level 1 for the next tile (other buffer)
			__syncwarp();
			if (lane == 0) mbar_arrive(&bar_e1[b]);
			if (t >= 2) mbar_wait(&bar_e3[b], (unsigned)(((t >> 1) + 1) & 1)); // warp 2 is done with L3(b) of tile t-2
			const int n3 = len >> 3;
			for (int j0 = lane * 5; j0 < n3; j0 += 160) ds2_run<5>(sm, L2, L3(b), j0);
			__syncwarp();
			ws_carry(sm, L2, L2, n2, lane);
			__syncwarp();
			if (lane == 0) mbar_arrive(&bar_f3[b]);
		}
	}
	else {
		// ---------------- level 3 -> levels 4..K -> 96 kHz -> 48 kHz -> HBM ----------------
		const float2 *rot_g = p.rot + (p.P >> K) + (base >> K);
		float2 *Cg = p.C + (long long)(stream * 2) * p.c_stride + p.c_off + (base >> (K + 1));
		const int m_first = p.P >> (K + 1);
		const int off_up = p.off_up + FE_HIST, off_dn = p.off_dn + FE_HIST, off_wa = p.off_wa + FE_HIST, off_wb = p.off_wb + FE_HIST;
		for (int t = 0; t < n_tiles; t++) {
			const int rel = t * p.tile;
			const int len = min(p.tile, span - rel);
			const int b = t & 1;
			mbar_wait(&bar_f3[b], (unsigned)((t >> 1) & 1));
			int src = L3(b);
#pragma unroll
			for (int l = 3; l < K; l++) {
				const int dst = p.off_lv[l + 1] + FE_HIST;
				const int n_out = len >> (l + 1);
				for (int j0 = lane * 5; j0 < n_out; j0 += 160) ds2_run<5>(sm, src, dst, j0);
				__syncwarp();
				if (l == 3) {
					ws_carry(sm, L3(b), L3(b ^ 1), len >> 3, lane);
					__syncwarp();
					if (lane == 0) mbar_arrive(&bar_e3[b]);
				}
				else ws_carry(sm, src, src, len >> l, lane);
				src = dst;
			}
			const int n96 = len >> K;
			const float2 *rg = rot_g + (rel >> K);
			for (int i = lane; i < n96; i += 32) {
				float2 x = sm[src + i];
				if (p.use_fdc) { // alpha * (h1 + data[i]) + h2 * beta
					const float2 tt = cadd(sm[src + i - 2], x);
					const float2 h2 = sm[src + i - 1];
					x = make_float2(__fadd_rn(__fmul_rn(p.fdc_alpha, tt.x), __fmul_rn(h2.x, p.fdc_beta)),
									__fadd_rn(__fmul_rn(p.fdc_alpha, tt.y), __fmul_rn(h2.y, p.fdc_beta)));
				}
				const float2 r = __ldg(rg + i);
				const float RR = __fmul_rn(x.x, r.x), II = __fmul_rn(x.y, r.y), RI = __fmul_rn(x.x, r.y), IR = __fmul_rn(x.y, r.x);
				sm[off_up + i] = make_float2(__fsub_rn(RR, II), __fadd_rn(IR, RI));
				sm[off_dn + i] = make_float2(__fadd_rn(RR, II), __fsub_rn(IR, RI));
			}
			__syncwarp();
			if (K == 3) { // level 3 fed the 96 kHz stage directly
				ws_carry(sm, L3(b), L3(b ^ 1), n96, lane);
				__syncwarp();
				if (lane == 0) mbar_arrive(&bar_e3[b]);
			}
			else ws_carry(sm, src, src, n96, lane);
			const int n48 = n96 >> 1;
			const int runs = (n48 + 4) / 5;
			for (int r = lane; r < 2 * runs; r += 32) {
				const int ch = r >= runs;
				ds2_run<5>(sm, ch ? off_dn : off_up, ch ? off_wb : off_wa, (ch ? r - runs : r) * 5);
			}
			__syncwarp();
			ws_carry(sm, off_up, off_up, n96, lane);
			ws_carry(sm, off_dn, off_dn, n96, lane);
			const int m_rel = rel >> (K + 1);
			if (m_rel + n48 > m_first) {
				const int m_lo = m_first - m_rel;
				for (int r = lane; r < 2 * runs; r += 32) {
					const int ch = r >= runs;
					fcic_run<5>(sm, ch ? off_wb : off_wa, Cg + (ch ? p.c_stride : 0) + m_rel, (ch ? r - runs : r) * 5, m_lo, n48);
				}
			}
			__syncwarp();
			ws_carry(sm, off_wa, off_wa, n48, lane);
			ws_carry(sm, off_wb, off_wb, n48, lane);
			__syncwarp();
		}
	}
}
#undef L1
#undef L3

// ---------------------------------------------------------------------------------------------
// K1'': the front end as the reference writes it -- a per-sample streaming pipeline with its state in registers --
// run by every THREAD on its own sub-segment of a stream.  A lane walks [a - P, a + S): the first P samples only warm
// the state up from zero (each CIC stage is a pure function of its last six inputs, see k_frontend), after that every
// 2^(K+1) inputs yield one 48 kHz sample per channel.  Per input pair a Downsample2CIC5 stage costs 9 packed adds and
// one packed multiply (DSP.cpp:93-117, literally: r_k = z; z += h_k / h_k = z; z += r_k) and nothing goes through shared
// memory between stages (ptxas fuses the exact 1/32 scaling of a stage with the first add of the next one into FFMA2;
// a power-of-two factor makes that bit-identical to the separate multiply unless a value is subnormal); shared memory only stages the input: the warp fetches the next 16 samples of all 32 lanes with
// coalesced 16-byte cp.async copies (raw format, converted when read) into a ring, four chunks ahead.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
struct Cic5 { c64 h0, h1, h2, h3, h4; };
__device__ __forceinline__ void cic5_zero(Cic5 &s) { s.h0 = s.h1 = s.h2 = s.h3 = s.h4 = 0ull; }
// one even/odd input pair of Downsample2CIC5 -> one output
__device__ __forceinline__ c64 ds2_pair(Cic5 &s, c64 xe, c64 xo, c64 sc) {
	c64 z = xe;
	const c64 r0 = z; z = padd(z, s.h0);
	const c64 r1 = z; z = padd(z, s.h1);
	const c64 r2 = z; z = padd(z, s.h2);
	const c64 r3 = z; z = padd(z, s.h3);
	const c64 r4 = z; z = padd(z, s.h4);
	const c64 out = pmul(z, sc);
	z = xo;
	s.h0 = z; z = padd(z, r0);
	s.h1 = z; z = padd(z, r1);
	s.h2 = z; z = padd(z, r2);
	s.h3 = z; z = padd(z, r3);
	s.h4 = z;
	(void)r4;
	return out;
}
// one even/odd input pair of FilterCIC5 -> two outputs (DSP.cpp:132-157)
__device__ __forceinline__ void fcic_pair(Cic5 &s, c64 xe, c64 xo, c64 sc, c64 &oe, c64 &oo) {
	c64 z = xe;
	const c64 r0 = z; z = padd(z, s.h0);
	const c64 r1 = z; z = padd(z, s.h1);
	const c64 r2 = z; z = padd(z, s.h2);
	const c64 r3 = z; z = padd(z, s.h3);
	const c64 r4 = z; z = padd(z, s.h4);
	oe = pmul(z, sc);
	z = xo;
	s.h0 = z; z = padd(z, r0);
	s.h1 = z; z = padd(z, r1);
	s.h2 = z; z = padd(z, r2);
	s.h3 = z; z = padd(z, r3);
	s.h4 = z; z = padd(z, r4);
	oo = pmul(z, sc);
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

constexpr int ST_WARPS = 4;  // warps per CTA (independent of each other)
constexpr int ST_G = 16;     // samples per lane per staged chunk
template <int FMT>
struct StFmt {
	static constexpr int BPS = FMT == 0 ? 8 : (FMT == 3 ? 4 : 2);
	static constexpr int CHUNK = ST_G * BPS;      // bytes of one lane's chunk: 128 / 32 / 32 / 64
	static constexpr int PIECES = CHUNK / 16;     // 16-byte pieces per lane chunk = cp.async instructions per warp chunk
	static constexpr int SLOT = CHUNK + 16;       // lane stride in the ring (odd multiple of 16 bytes: conflict-free 16-byte reads)
};
// sample pair j (samples 2j, 2j+1) of a staged chunk
template <int FMT>
__device__ __forceinline__ void st_read_pair(const unsigned char *slot, int j, c64 &xe, c64 &xo) {
	if (FMT == 0) {
		const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(slot + j * 16);
		xe = v.x;
		xo = v.y;
	}
	else if (FMT == 1) {
		const uchar4 v = *reinterpret_cast<const uchar4 *>(slot + j * 4);
		xe = pack2(__fmul_rn((float)((int)v.x - 128), 0.0078125f), __fmul_rn((float)((int)v.y - 128), 0.0078125f));
		xo = pack2(__fmul_rn((float)((int)v.z - 128), 0.0078125f), __fmul_rn((float)((int)v.w - 128), 0.0078125f));
	}
	else if (FMT == 2) {
		const char4 v = *reinterpret_cast<const char4 *>(slot + j * 4);
		xe = pack2(__fmul_rn((float)v.x, 0.0078125f), __fmul_rn((float)v.y, 0.0078125f));
		xo = pack2(__fmul_rn((float)v.z, 0.0078125f), __fmul_rn((float)v.w, 0.0078125f));
	}
	else {
		const short4 v = *reinterpret_cast<const short4 *>(slot + j * 8);
		xe = pack2(__fmul_rn((float)v.x, 3.0517578125e-05f), __fmul_rn((float)v.y, 3.0517578125e-05f));
		xo = pack2(__fmul_rn((float)v.z, 3.0517578125e-05f), __fmul_rn((float)v.w, 3.0517578125e-05f));
	}
}

template <int FMT, int K, int ST_NB, bool PRE = false>
__global__ void __launch_bounds__(ST_WARPS * 32) k_frontend_st(const FeParams p) {
	static_assert(K >= 3 && K <= 7, "streaming front end: 768 kS/s .. 12288 kS/s");
	typedef StFmt<FMT> F;
	constexpr int SS = 1 << (K + 2);     // inputs per super-step: two 48 kHz samples per channel
	constexpr int NCH = SS / ST_G;       // chunks per super-step
	constexpr int N96 = SS >> K;         // 96 kHz samples per super-step (4)
	extern __shared__ __align__(16) unsigned char st_ring[]; // [ST_WARPS][ST_NB][32 * SLOT]
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const long long wg = (long long)blockIdx.x * ST_WARPS + wib;
	const int stream = (int)(wg / p.st_wps);
	if (stream >= p.st_B) return; // whole warp
	const int sub0 = (int)(wg - (long long)stream * p.st_wps) * 32;
	const int S = p.st_S;
	unsigned char(*ring)[32 * F::SLOT] = reinterpret_cast<unsigned char(*)[32 * F::SLOT]>(st_ring + (size_t)wib * ST_NB * 32 * F::SLOT);
	// every lane owns S samples [a, a + S) (the host only picks this kernel when N is a multiple of 32 * S * st_wps)
	const long long a = (long long)(sub0 + lane) * S;
	const int warp_chunks = (S + p.P) / ST_G; // warm-up included
	const unsigned char *in_b = reinterpret_cast<const unsigned char *>(p.in) + (long long)stream * p.in_stride * F::BPS;
	const unsigned char *tl_b = reinterpret_cast<const unsigned char *>(p.tail) + ((long long)stream * p.P + p.P) * F::BPS;
	// staging: in instruction `it` lane j fetches 16-byte piece (j % PIECES) of the chunk of owner it*(32/PIECES) + j / PIECES
	constexpr int OWN_PER_IT = 32 / F::PIECES;
	const int o0 = lane / F::PIECES, q0 = lane % F::PIECES;
	const long long lane_off = ((long long)(sub0 + o0) * S - p.P) * F::BPS + q0 * 16; // byte offset of chunk 0, piece q0, owner o0
	const long long it_step = (long long)OWN_PER_IT * S * F::BPS;                        // next instruction: next group of owners
	const int dst_off = o0 * F::SLOT + q0 * 16;
	const bool from_tail = sub0 == 0 && o0 == 0; // only the first sub-segment of a stream starts in the previous submit
	auto prefetch = [&](int c) {
		if (c < warp_chunks) {
			const long long coff = lane_off + (long long)c * (ST_G * F::BPS);
			unsigned char *dst = &ring[c % ST_NB][dst_off];
#pragma unroll
			for (int it = 0; it < F::PIECES; it++) {
				const unsigned char *src = ((it == 0 && from_tail && c * ST_G < p.P) ? tl_b : in_b) + coff + it * it_step;
				cp_async16(dst + it * OWN_PER_IT * F::SLOT, src);
			}
		}
		cp_async_commit();
	};
	const c64 sc = pack2(0.03125f, 0.03125f);
	Cic5 lv[K], chA, chB, fA, fB;
#pragma unroll
	for (int l = 0; l < K; l++) cic5_zero(lv[l]);
	cic5_zero(chA); cic5_zero(chB); cic5_zero(fA); cic5_zero(fB);
	c64 fd1 = 0ull, fd2 = 0ull; // FilterComplex3Tap h1, h2
	// PRE: decimation in front of DSP::Upsample -- the level-K samples go to D0 and nothing else is computed
	const float2 *rot_g = PRE ? nullptr : p.rot + (p.P >> K) + ((a - p.P) >> K);
	float2 *Cg = PRE ? p.D0 + (long long)stream * p.d0_stride + p.d0_off + ((a - p.P) >> K)
					 : p.C + (long long)(stream * 2) * p.c_stride + p.c_off + ((a - p.P) >> (K + 1));
	const int n_super = warp_chunks / NCH;
	const int warm_super = p.P / SS;
#pragma unroll
	for (int c = 0; c < ST_NB - 1; c++) prefetch(c);
	// Rotate phasors: loaded two super-steps ahead of their use (under load a global load can take longer than one
	// super-step of arithmetic)
	float2 rt_n1[N96], rt_n2[N96];
#pragma unroll
	for (int i = 0; i < N96; i++) {
		rt_n1[i] = PRE ? make_float2(0.f, 0.f) : __ldg(rot_g + i);
		rt_n2[i] = PRE ? make_float2(0.f, 0.f) : __ldg(rot_g + (n_super > 1 ? N96 : 0) + i);
	}
	for (int ss = 0; ss < n_super; ss++) {
		float2 rt[N96];
#pragma unroll
		for (int i = 0; i < N96; i++) {
			rt[i] = rt_n1[i];
			rt_n1[i] = rt_n2[i];
		}
		if (!PRE && ss + 2 < n_super) {
#pragma unroll
			for (int i = 0; i < N96; i++) rt_n2[i] = __ldg(rot_g + (ss + 2) * N96 + i);
		}
		c64 lvK[N96]; // PRE: the super-step's level-K outputs
		c64 pend[K + 1];  // pend[l]: even-indexed input waiting at level l+1 (l = 1..K-1), pend[K]: unused
		c64 upE = 0ull, dnE = 0ull, waE = 0ull, wbE = 0ull;
		c64 outA0 = 0ull, outA1 = 0ull, outB0 = 0ull, outB1 = 0ull;
#pragma unroll
		for (int cc = 0; cc < NCH; cc++) {
			const int c = ss * NCH + cc;
			prefetch(c + ST_NB - 1);
			cp_async_wait<ST_NB - 1>(); // chunk c has landed
			__syncwarp();
			{
				const unsigned char *slot = &ring[c % ST_NB][lane * F::SLOT];
#pragma unroll
				for (int j = 0; j < ST_G / 2; j++) {
					const int n1 = cc * (ST_G / 2) + j; // index of this pair's output at level 1 within the super-step
					c64 xe, xo;
					st_read_pair<FMT>(slot, j, xe, xo);
					c64 y = ds2_pair(lv[0], xe, xo, sc);
					// ripple through the deeper levels: an output with an odd index completes a pair one level down
					int idx = n1;
					bool live = true;
#pragma unroll
					for (int l = 1; l < K; l++) {
						if (live) {
							if ((idx & 1) == 0) { pend[l] = y; live = false; }
							else { y = ds2_pair(lv[l], pend[l], y, sc); idx >>= 1; }
						}
					}
					if (live && PRE) lvK[idx] = y;
					if (live && !PRE) { // y is 96 kHz sample idx (0..N96-1) of the super-step
						c64 x = y;
						if (p.use_fdc) { // FilterComplex3Tap: alpha * (h1 + x) + h2 * beta (DSP.cpp:283-293)
							// scalar intrinsics: ptxas would contract a packed mul + add pair into FFMA2 here, and these products are not exact
							const float2 h1 = unpack2(fd1), h2 = unpack2(fd2), yv = unpack2(y);
							const float tx = __fadd_rn(h1.x, yv.x), ty = __fadd_rn(h1.y, yv.y);
							x = pack2(__fadd_rn(__fmul_rn(p.fdc_alpha, tx), __fmul_rn(h2.x, p.fdc_beta)),
									  __fadd_rn(__fmul_rn(p.fdc_alpha, ty), __fmul_rn(h2.y, p.fdc_beta)));
							fd1 = fd2;
							fd2 = y;
						}
						const float2 xv = unpack2(x);
						const float2 r = rt[idx];
						const float RR = __fmul_rn(xv.x, r.x), II = __fmul_rn(xv.y, r.y), RI = __fmul_rn(xv.x, r.y), IR = __fmul_rn(xv.y, r.x);
						const c64 up = pack2(__fsub_rn(RR, II), __fadd_rn(IR, RI));
						const c64 dn = pack2(__fadd_rn(RR, II), __fsub_rn(IR, RI));
						if ((idx & 1) == 0) { upE = up; dnE = dn; }
						else {
							const c64 wa = ds2_pair(chA, upE, up, sc), wb = ds2_pair(chB, dnE, dn, sc);
							if ((idx & 2) == 0) { waE = wa; wbE = wb; }
							else {
								fcic_pair(fA, waE, wa, sc, outA0, outA1);
								fcic_pair(fB, wbE, wb, sc, outB0, outB1);
							}
						}
					}
				}
			}
			__syncwarp(); // the ring slot may be refilled by a later prefetch
		}
		if (PRE) {
			if (ss >= warm_super) {
				float2 *o = Cg + ss * N96;
#pragma unroll
				for (int i = 0; i < N96; i += 2) *reinterpret_cast<ulonglong2 *>(o + i) = make_ulonglong2(lvK[i], lvK[i + 1]);
			}
		}
		else if (ss >= warm_super) { // two 48 kHz samples per channel
			float2 *o = Cg + ss * 2;
			*reinterpret_cast<ulonglong2 *>(o) = make_ulonglong2(outA0, outA1);
			*reinterpret_cast<ulonglong2 *>(o + p.c_stride) = make_ulonglong2(outB0, outB1);
		}
	}
	cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// DSP::Upsample (DSP.cpp:192-212): out = (1 - alpha) * a + alpha * b with a, b consecutive inputs.  alpha is a float
// accumulator that only depends on how many samples have gone by, so the host replays it (same binary32 adds) and
// hands the kernel one (input index, alpha) pair per output; outputs land in a ring of whole reference blocks.
// ---------------------------------------------------------------------------------------------
__global__ void k_upsample(const float2 *__restrict__ D0, long long d0_stride, int d0_off, const int *__restrict__ src, const float *__restrict__ alpha,
						   int M, float2 *__restrict__ S, long long s_stride, long long m0, int cap) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= M) return;
	const float2 *d = D0 + (long long)blockIdx.y * d0_stride + d0_off;
	const int i = src[m];
	const float al = alpha[m];
	const float2 a = d[i - 1], b = d[i];
	const float w = __fsub_rn(1.0f, al);
	float2 o;
	o.x = __fadd_rn(__fmul_rn(w, a.x), __fmul_rn(al, b.x));
	o.y = __fadd_rn(__fmul_rn(w, a.y), __fmul_rn(al, b.y));
	S[(long long)blockIdx.y * s_stride + (int)((m0 + m) % cap)] = o;
}

// DSP::DownsampleKFilter with Filters::BlackmanHarris_28_3, K = 3 (DSP.cpp:160-189, Filters.h:43-53; the 288 kS/s
// front end, Model.cpp:308-313): out[j] = sum_k taps[k] * x[n_j - 25 + k], n_j = first + 3 j, accumulated from 0 in
// ascending k.  Input: the submit's samples, negative indices from the previous submit's tail.
constexpr int DSK_T = 26;
constexpr int DSK_THREADS = 256;
__constant__ float c_taps_bh28_3[DSK_T];
template <int FMT>
__device__ __forceinline__ float2 fe_load_one(const void *base, long long idx) {
	if (FMT == 0) return __ldg(reinterpret_cast<const float2 *>(base) + idx);
	if (FMT == 1) {
		const uchar2 v = __ldg(reinterpret_cast<const uchar2 *>(base) + idx);
		return make_float2(__fmul_rn((float)((int)v.x - 128), 0.0078125f), __fmul_rn((float)((int)v.y - 128), 0.0078125f));
	}
	if (FMT == 2) {
		const char2 v = __ldg(reinterpret_cast<const char2 *>(base) + idx);
		return make_float2(__fmul_rn((float)v.x, 0.0078125f), __fmul_rn((float)v.y, 0.0078125f));
	}
	const short2 v = __ldg(reinterpret_cast<const short2 *>(base) + idx);
	return make_float2(__fmul_rn((float)v.x, 3.0517578125e-05f), __fmul_rn((float)v.y, 3.0517578125e-05f));
}
template <int FMT>
__global__ void __launch_bounds__(DSK_THREADS) k_dsk(const void *__restrict__ in, long long in_stride, const void *__restrict__ tail, int tail_len, int first,
													  int n_out, float2 *__restrict__ S, long long s_stride, long long j0, int cap) {
	__shared__ float2 x[3 * DSK_THREADS + DSK_T];
	const int stream = blockIdx.y, tid = threadIdx.x;
	const int o0 = blockIdx.x * DSK_THREADS;             // first output of this CTA
	const int lo = first + 3 * o0 - (DSK_T - 1);         // input index of x[0], relative to the submit
	for (int i = tid; i < 3 * DSK_THREADS + DSK_T; i += DSK_THREADS) {
		const int n = lo + i;
		float2 v = make_float2(0.f, 0.f);
		if (n < 0) {
			if (n >= -tail_len) v = fe_load_one<FMT>(tail, (long long)stream * tail_len + tail_len + n);
		}
		else if (n <= first + 3 * (n_out - 1)) v = fe_load_one<FMT>(in, (long long)stream * in_stride + n);
		x[i] = v;
	}
	__syncthreads();
	const int o = o0 + tid;
	if (o >= n_out) return;
	float2 acc = make_float2(0.f, 0.f);
#pragma unroll
	for (int k = 0; k < DSK_T; k++) {
		const float2 dd = x[3 * tid + k];
		acc.x = __fadd_rn(acc.x, __fmul_rn(c_taps_bh28_3[k], dd.x));
		acc.y = __fadd_rn(acc.y, __fmul_rn(c_taps_bh28_3[k], dd.y));
	}
	S[(long long)stream * s_stride + (int)((j0 + o) % cap)] = acc;
}

// new_tail = last P samples of (old_tail ++ chunk); works for any N.  Copies 8-byte words (P is a multiple of 4
// samples and every format has >= 2 bytes per sample, so rows and offsets stay 8-byte aligned).
__global__ void k_tail_update(uint2 *__restrict__ new_tail, const uint2 *__restrict__ old_tail, const uint2 *__restrict__ in,
							  long long in_stride_w, long long n_w, int p_w) {
	const int stream = blockIdx.y;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p_w; i += gridDim.x * blockDim.x) {
		const long long s = (long long)i + n_w - p_w; // word index relative to chunk start
		new_tail[(long long)stream * p_w + i] = s >= 0 ? in[(long long)stream * in_stride_w + s] : old_tail[(long long)stream * p_w + (s + p_w)];
	}
}

// ---------------------------------------------------------------------------------------------
// small utility: move `cnt` trailing elements of each row to the slot just before `dst_end`
// (keeps unconsumed samples / filter history in front of the next submit's data)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_carry(T *__restrict__ buf, long long stride, int src_begin, int dst_begin, int cnt) {
	extern __shared__ __align__(16) unsigned char carry_sm[];
	T *tmp = reinterpret_cast<T *>(carry_sm);
	T *row = buf + (long long)blockIdx.x * stride;
	for (int i = threadIdx.x; i < cnt; i += blockDim.x) tmp[i] = row[src_begin + i];
	__syncthreads();
	for (int i = threadIdx.x; i < cnt; i += blockDim.x) row[dst_begin + i] = tmp[i];
}

// copy `cnt` elements of each row from one buffer to another (unconsumed samples / filter history handed to the
// buffer the next submit's front end writes into)
template <typename T>
__global__ void k_carry2(const T *__restrict__ src, T *__restrict__ dst, long long stride, int src_begin, int dst_begin, int cnt) {
	const T *srow = src + (long long)blockIdx.x * stride + src_begin;
	T *drow = dst + (long long)blockIdx.x * stride + dst_begin;
	for (int i = threadIdx.x; i < cnt; i += blockDim.x) drow[i] = srow[i];
}

// ---------------------------------------------------------------------------------------------
// K2a: SquareFreqOffsetCorrection, estimation half (DSP.cpp:417-455, FFT.h:93-130).
// One warp per 512-sample block: x^2 in bit-reversed order, the reference's radix-2 DIT butterflies stage by
// stage, |F| in fftshift order; then (one lane per block) the sequential float cumsum, then the parallel
// first-maximum searches.  Result: an index into the host-built phasor-step table.
// ---------------------------------------------------------------------------------------------
constexpr int CGF_N = 512;
constexpr int CGF_BLK_PER_CTA = 16;
constexpr int CGF_THREADS = 256;
constexpr int CGF_ROWP = 513;           // padded row (floats) so 16 lanes scanning 16 rows hit 16 banks
constexpr int CGF_IDX_OFFSET = 3;       // idx = i + 3, i in [-3, 410]
constexpr int CGF_IDX_NONE = 414;       // no bin above zero: fz = -1
constexpr int CGF_NIDX = 415;

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
constexpr int FIRC_T = 17;
constexpr int FIRC_TILE = 256;
__constant__ float c_taps_coherent[FIRC_T];
__constant__ float c_taps_receiver[37];

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

// ---------------------------------------------------------------------------------------------
// K2-FM: Demod::FM (Demod.cpp:27-37) fused with DSP::Filter 37 taps (DSP.cpp:249-280, Filters.h:24-33).
// Cbuf keeps 37 samples of history in front of the new ones, so the FM values feeding the FIR history are
// recomputed instead of stored.
// ---------------------------------------------------------------------------------------------
constexpr int FIRF_T = 37;
constexpr int FIRF_TILE = 256;
__global__ void __launch_bounds__(FIRF_TILE) k_fm_fir(const float2 *__restrict__ Cbuf, long long c_stride, int c_new, int n,
														 float *__restrict__ Fbuf, long long f_stride, int f_off,
														 float *__restrict__ tap_fm, long long tap_stride) {
	__shared__ float fm[FIRF_TILE + FIRF_T - 1];
	const int row = blockIdx.y, t0 = blockIdx.x * FIRF_TILE, tid = threadIdx.x;
	const float2 *c = Cbuf + (long long)row * c_stride + c_new; // c[i] = new sample i, history at negative i
	for (int i = tid; i < FIRF_TILE + FIRF_T - 1; i += FIRF_TILE) {
		const int m = t0 + i - (FIRF_T - 1);
		float v = 0.0f;
		if (m < n) {
			const float2 a = c[m], pv = c[m - 1];
			// data[i] * conj(prev): re = a.re*p.re - a.im*(-p.im), im = a.re*(-p.im) + a.im*p.re
			const float re = __fsub_rn(__fmul_rn(a.x, pv.x), __fmul_rn(a.y, -pv.y));
			const float im = __fadd_rn(__fmul_rn(a.x, -pv.y), __fmul_rn(a.y, pv.x));
			v = __fdiv_rn(fd_atan2f(im, re), 3.14159265358979323846f);
			if (tap_fm && m >= 0 && i >= FIRF_T - 1) tap_fm[(long long)row * tap_stride + m] = v;
		}
		fm[i] = v;
	}
	__syncthreads();
	const int m = t0 + tid;
	if (m < n) {
		float x = 0.0f;
#pragma unroll
		for (int k = 0; k < FIRF_T; k++) x = __fadd_rn(x, __fmul_rn(c_taps_receiver[k], fm[tid + k]));
		Fbuf[(long long)row * f_stride + f_off + m] = x;
	}
}

// K2-FM': the same FM + FIR37, five outputs (one symbol slot of the 5-phase deinterleaver, DSP.h:65-73) per thread:
// 41 discriminator values are read once into registers and reused by the five 37-tap sums (each still accumulated
// in the reference's order, k = 0..36 from 0.0f).  Besides the filtered samples the kernel emits what the decoders
// actually consume: one sign bit per (row, sampling phase, slot), packed 32 slots per word by warp ballots.
constexpr int FM5_THREADS = 128; // slots per CTA
constexpr int FM5_SAMPLES = FM5_THREADS * 5;
struct Fm5Params {
	const float2 *Cbuf;
	long long c_stride;
	int c_new, n;        // new samples start at Cbuf[row][c_new], n of them
	int r0;              // abs index of new sample 0 modulo 5: slot 0 starts r0 samples before it
	int nslots;
	float *Fbuf;         // FIR37 output, [rows][f_stride], new sample m at f_off + m
	long long f_stride;
	int f_off;
	uint32_t *dbits;     // [rows*5][dwords]
	int dwords;
	float *tap_fm;       // optional
	long long tap_stride;
	float *tap_dec;      // optional: decoder input samples [rows*5][nslots], valid ones only, packed per phase
};
__global__ void __launch_bounds__(FM5_THREADS) k_fm_fir5(const Fm5Params p) {
	__shared__ float fm[FM5_SAMPLES + FIRF_T - 1 + 3];
	const int row = blockIdx.y, tid = threadIdx.x;
	const int S0 = blockIdx.x * FM5_THREADS;
	const int M0 = 5 * S0 - p.r0; // new-sample index of the first sample of slot S0
	const float2 *c = p.Cbuf + (long long)row * p.c_stride + p.c_new;
	for (int i = tid; i < FM5_SAMPLES + FIRF_T - 1; i += FM5_THREADS) {
		const int m = M0 + i - (FIRF_T - 1);
		float v = 0.0f;
		if (m < p.n && m >= -(FIRF_T - 1) - 4) {
			const float2 a = c[m], pv = c[m - 1];
			const float re = __fsub_rn(__fmul_rn(a.x, pv.x), __fmul_rn(a.y, -pv.y));
			const float im = __fadd_rn(__fmul_rn(a.x, -pv.y), __fmul_rn(a.y, pv.x));
			v = __fdiv_rn(fd_atan2f(im, re), 3.14159265358979323846f);
			if (p.tap_fm && m >= 0 && i >= FIRF_T - 1) p.tap_fm[(long long)row * p.tap_stride + m] = v;
		}
		fm[i] = v;
	}
	__syncthreads();
	float x[FIRF_T + 4];
#pragma unroll
	for (int i = 0; i < FIRF_T + 4; i++) x[i] = fm[5 * tid + i];
	float y[5];
#pragma unroll
	for (int j = 0; j < 5; j++) {
		float acc = 0.0f;
#pragma unroll
		for (int k = 0; k < FIRF_T; k++) acc = __fadd_rn(acc, __fmul_rn(c_taps_receiver[k], x[j + k]));
		y[j] = acc;
	}
	const int slot = S0 + tid;
	const int m0 = M0 + 5 * tid;
#pragma unroll
	for (int j = 0; j < 5; j++) {
		const int m = m0 + j;
		if (m >= 0 && m < p.n) {
			p.Fbuf[(long long)row * p.f_stride + p.f_off + m] = y[j];
			if (p.tap_dec) p.tap_dec[(long long)(row * 5 + j) * p.nslots + slot - (j >= p.r0 ? 0 : 1)] = y[j];
		}
		const unsigned w = __ballot_sync(0xffffffffu, y[j] > 0.0f);
		if ((tid & 31) == j && (slot >> 5) < p.dwords) p.dbits[(long long)(row * 5 + j) * p.dwords + (slot >> 5)] = w;
	}
}

// ---------------------------------------------------------------------------------------------
// K3: symbol timing + demodulation + bit decoder.
//   ModelDefault : ScatterPLL (DSP.h:95-117) -> 5 x PhaseSearchEMA / PhaseSearch (Demod.cpp:39-170) -> 5 x Decoder
//   ModelStandard: Deinterleave (DSP.h:65-73) -> 5 x Decoder
//   ModelBase    : SimplePLL (DSP.cpp:28-57) -> 1 x Decoder
// One thread per (row, sampling phase); the five phases of a row sit in five adjacent lanes of one warp so the
// decoder's Reset broadcast (AIS.cpp:47-49, Model.cpp:566-573) is a warp vote.  Frame bits live in shared memory.
// ---------------------------------------------------------------------------------------------
enum { ST_TRAINING = 0, ST_STARTFLAG = 1, ST_DATAFCS = 3 };
constexpr int DEC_WORDS = 35;     // 140 bytes (Message.h:69 data[MAX_AIS_FRAME_BYTES + 4])
constexpr int MAX_FRAME_BITS = 1087; // MAX_AIS_FRAME_LENGTH (Message.h:41)
constexpr int K3_THREADS = 32; // one warp per CTA: the rows are few, spread them over all SMs

struct DecState { // one per (row, phase); persisted between submits (frame bits live in a separate array)
	int state, lastBit, prev, position, one_seq;
	float level;
	long long start_idx;
};
struct PsState { // PhaseSearchEMA (Demod.h:68-86) / PhaseSearch (Demod.h:41-66)
	float ma[16];
	uint32_t plane[5]; // plane[d] bit h = decision of hypothesis h, d symbols ago (bits[h] >> d & 1)
	int max_idx, rot, last;
};
struct FrameRec {
	int row, phase, nbits;
	float level;          // TAG::level before the dB conversion (AIS.h:147)
	float ppm;
	int chunk;            // ordinal of the submit
	long long start_idx, end_idx;
	uint32_t data[DEC_WORDS];
	int blk;              // ordinal of the front-end block (several per submit behind a resampler)
};

struct DecCtx {
	uint32_t *frame; // shared memory, word w of this thread at frame[w * K3_THREADS]
	int mode_level;
};

__device__ __forceinline__ uint32_t frame_word(const DecCtx &c, int w) { return c.frame[w * K3_THREADS]; }
__device__ __forceinline__ int dec_type(const DecCtx &c) { return (frame_word(c, 0) & 0xff) >> 2; }
__device__ __forceinline__ unsigned dec_mmsi(const DecCtx &c) {
	const uint32_t w0 = frame_word(c, 0), w1 = frame_word(c, 1);
	const unsigned d1 = (w0 >> 8) & 0xff, d2 = (w0 >> 16) & 0xff, d3 = (w0 >> 24) & 0xff, d4 = w1 & 0xff;
	return (d1 << 22) | (d2 << 14) | (d3 << 6) | (d4 >> 2);
}
__device__ __forceinline__ bool dec_cannot_be_valid(const DecCtx &c, int len) { // AIS.cpp:111-142
	if (len < 30) return false;
	const int t = dec_type(c);
	switch (len) {
	case 30: return t > 28 || t == 0;
	case 62: return dec_mmsi(c) > 999999999u;
	case 96: return t == 10;
	case 168: return t == 16;
	case 184: return t == 15 || t == 20 || t == 23;
	case 192: return t == 1 || t == 2 || t == 3 || t == 4 || t == 7 || t == 9 || t == 11 || t == 18 || t == 22 || t == 24 || t == 25 || t == 27 || t == 28;
	case 336: return t == 19;
	case 385: return t == 21;
	case 448: return t == 5;
	}
	return false;
}
// Same CRC (AIS.cpp:55-64: reflected 0x8408, init 0xFFFF, good residue 0xF0B8), eight bits per step: the frame words
// hold the bits LSB first, which is the order the reflected CRC consumes them.
__device__ __forceinline__ bool dec_crc16_bytes(const DecCtx &c, int len) {
	unsigned crc = 0xFFFF;
	const int nbytes = len >> 3;
	uint32_t w = 0;
	for (int k = 0; k < nbytes; k++) {
		if ((k & 3) == 0) w = frame_word(c, k >> 2);
		unsigned dta = ((w >> ((k & 3) * 8)) ^ crc) & 0xffu;
		dta ^= (dta << 4) & 0xffu;
		crc = (((dta << 8) | (crc >> 8)) ^ (dta >> 4) ^ (dta << 3)) & 0xffffu;
	}
	for (int i = nbytes * 8; i < len; i++) {
		const unsigned bit = (frame_word(c, i >> 5) >> (i & 31)) & 1u;
		crc = ((bit ^ crc) & 1u) ? ((crc >> 1) ^ 0x8408u) : (crc >> 1);
	}
	return crc == 0xF0B8u;
}
__device__ __forceinline__ bool dec_crc16(const DecCtx &c, int len) { // AIS.cpp:55-64
	unsigned crc = 0xFFFF;
	for (int i = 0; i < len; i++) {
		const unsigned bit = (frame_word(c, i >> 5) >> (i & 31)) & 1u;
		crc = ((bit ^ crc) & 1u) ? ((crc >> 1) ^ 0x8408u) : (crc >> 1);
	}
	return crc == 0xF0B8u;
}

// One Decoder::Run (AIS.h:91-181).  Returns true when a frame with a good CRC just completed (processData true);
// in that case fr_len = payload bits + 16 and the caller emits and performs the FOUNDMESSAGE/Reset protocol.
__device__ __forceinline__ bool dec_step(DecState &d, const DecCtx &c, float sample, float sample_lvl, long long sample_idx, int &fr_len,
										 float &fr_level, int &lastBit_before) {
	const int dd = sample > 0.0f;
	const int Bit = !(dd ^ d.prev);
	d.prev = dd;
	lastBit_before = d.lastBit;
	bool found = false;
	switch (d.state) {
	case ST_TRAINING:
		if (Bit != d.lastBit) d.position++;
		else {
			if (d.position > 4) {
				d.start_idx = sample_idx;
				d.state = ST_STARTFLAG;
				d.position = Bit ? 3 : 1;
				d.one_seq = 0;
			}
			else { d.state = ST_TRAINING; d.position = 0; d.one_seq = 0; }
		}
		break;
	case ST_STARTFLAG:
		if (d.position == 7) {
			if (Bit == 0) {
				d.state = ST_DATAFCS; d.position = 0; d.one_seq = 0;
				d.level = 0.0f;
				for (int w = 0; w < DEC_WORDS; w++) c.frame[w * K3_THREADS] = 0u; // msg.clear()
			}
			else { d.state = ST_TRAINING; d.position = 0; d.one_seq = 0; }
		}
		else {
			if (Bit == 1) d.position++;
			else { d.state = ST_TRAINING; d.position = 0; d.one_seq = 0; }
		}
		break;
	case ST_DATAFCS: {
		const int pos = d.position++;
		if (pos < MAX_FRAME_BITS) { // Message::setBit (Message.h:264-273)
			uint32_t *wp = &c.frame[(pos >> 5) * K3_THREADS];
			const uint32_t m = 1u << (pos & 31);
			*wp = Bit ? (*wp | m) : (*wp & ~m);
		}
		if (c.mode_level) d.level = __fadd_rn(d.level, sample_lvl);
		if (Bit == 1) {
			if (d.one_seq == 5) {
				fr_level = c.mode_level ? __fdiv_rn(d.level, (float)d.position) : 0.0f;
				const int len = d.position - 7;
				if (len >= 16 && dec_crc16(c, len)) {
					found = true;
					fr_len = len;
				}
				d.state = ST_TRAINING; d.position = 0; d.one_seq = 0;
			}
			else d.one_seq++;
		}
		else {
			if (d.one_seq == 5) d.position--;
			d.one_seq = 0;
		}
		if (d.position == MAX_FRAME_BITS || dec_cannot_be_valid(c, d.position)) { d.state = ST_TRAINING; d.position = 0; d.one_seq = 0; }
		break;
	}
	default: break;
	}
	d.lastBit = Bit;
	return found;
}

__device__ __forceinline__ void emit_frame(FrameRec *__restrict__ ring, int *__restrict__ ring_count, int ring_cap, int chunk, int blk, const DecCtx &c,
										   int row, int phase, int len, float level, float ppm, long long start_idx, long long end_idx) {
	const int slot = atomicAdd(ring_count, 1);
	if (slot >= ring_cap) return;
	FrameRec &r = ring[slot];
	r.row = row;
	r.phase = phase;
	r.nbits = len - 16;
	r.level = level;
	r.ppm = ppm;
	r.chunk = chunk;
	r.blk = blk;
	r.start_idx = start_idx;
	r.end_idx = end_idx;
	for (int w = 0; w < DEC_WORDS; w++) r.data[w] = frame_word(c, w);
}

__constant__ float c_ps_cos[8];
__constant__ float c_ps_sin[8];

struct K3Params {
	int ps_ema;
	int rows;
	int nsym;             // symbol slots (groups of 5 samples) to walk this submit
	long long e_stride;
	int e_begin;          // index in the row of the sample with absolute index abs_begin
	long long abs_begin;  // absolute per-channel index (TAG::sample_idx, DSP.h:110) of slot 0 / phase 0; multiple of 5
	long long abs_lo, abs_hi; // samples with abs_lo <= index < abs_hi exist this submit (Deinterleave forwards partial groups)
	const float2 *Ec;     // ModelDefault: FIR17 output
	const float *Ef;      // FM models: FIR37 output
	PsState *ps;
	float *ps_mem;        // PhaseSearch history |t| [16*12][rows*5] (only when !ps_ema)
	uint32_t *dbits;      // ModelDefault: demodulated bits, [rows*5][dwords], bit (s & 31) of word (s >> 5) = symbol s
	int dwords;
	float *lvl;           // ModelDefault: ScatterPLL level of symbol s (TAG::sample_lvl, DSP.h:100-106), [rows][lvl_stride]
	int lvl_stride;
	DecState *dec;
	uint32_t *dec_data;   // [DEC_WORDS][rows*5]
	FrameRec *ring;
	int *ring_count;
	int ring_cap;
	int chunk;
	int blk;
	int mode_level;
	// tag.ppm lookup (ModelDefault): block index of a sample = (abs_idx - blk_abs0) >> 9
	const int *stepidx;
	const float *ppmtab;
	long long blk_abs0;
	int nblk;
	float *tap_dec;       // optional: decoder input samples [rows*5][nsym]
	long long *dbg;       // optional per-row counters [rows][4]: cycles, frame-collecting steps, CRC runs, CRC bits
};

constexpr int K3_TS = 32;                 // symbols staged per tile
constexpr int K3_ROWLEN = K3_TS * 5;      // samples of one row in a tile

__device__ __forceinline__ void cp_async_f(float *smem_dst, const float *gsrc) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_f(float2 *smem_dst, const float2 *gsrc) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

// ---------------------------------------------------------------------------------------------
// K3a: PhaseSearchEMA / PhaseSearch (Demod.cpp:39-170), hypothesis-parallel.  Half a warp per (row, sampling
// phase): lane h owns hypothesis h (its EMA / 12-sample history and its last 5 sign decisions); the +-1 (+-2)
// neighbourhood argmax is three (five) shuffles.  The only thing leaving the kernel is one bit per symbol.
// ---------------------------------------------------------------------------------------------
constexpr int PS_THREADS = 128;
__global__ void __launch_bounds__(PS_THREADS) k_phase_search(const K3Params p) {
	__shared__ float2 tile[PS_THREADS / 32][2][2 * K3_ROWLEN];
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const int half = lane >> 4, h = lane & 15;
	const long long ninst = (long long)p.rows * 5;
	const long long warp_global = (long long)blockIdx.x * (PS_THREADS / 32) + wib;
	const long long inst = warp_global * 2 + half;
	const bool active = inst < ninst;
	const int row = active ? (int)(inst / 5) : 0, phase = active ? (int)(inst - (long long)row * 5) : 0;
	const unsigned hmask = 0xffffu << (half * 16);
	const int j = h < 8 ? h : 15 - h;
	const float cj = c_ps_cos[j];
	const float sj = h < 8 ? c_ps_sin[j] : -c_ps_sin[j]; // a - b == a + (-b) and im * (-s) == -(im * s), exactly
	const float weight = 0.85f, omw = __fsub_rn(1.0f, 0.85f);

	float ma = 0.0f, mem[12];
	uint32_t hist = 0; // bit d = sign decision of this hypothesis d symbols ago (uint8_t bits[] of the reference, low 5 bits)
	int max_idx = 0, rot = 0, last = 0;
#pragma unroll
	for (int l = 0; l < 12; l++) mem[l] = 0.0f;
	if (active) {
		const PsState &st = p.ps[inst];
		ma = st.ma[h];
#pragma unroll
		for (int dd = 0; dd < 5; dd++) hist |= ((st.plane[dd] >> h) & 1u) << dd;
		max_idx = st.max_idx;
		rot = st.rot;
		last = st.last;
		if (!p.ps_ema) {
#pragma unroll
			for (int l = 0; l < 12; l++) mem[l] = p.ps_mem[(long long)(h * 12 + l) * ninst + inst];
		}
	}
	const int nsamp = p.nsym * 5;
	float2(*mytile)[2 * K3_ROWLEN] = tile[wib];
	// each half stages the samples of its own row
	auto prefetch = [&](int buf, int s0) {
		const int base = s0 * 5;
		if (active) {
			const float2 *src = p.Ec + (long long)row * p.e_stride + p.e_begin + base;
			float2 *dst = &mytile[buf][half * K3_ROWLEN];
			for (int e = h; e < K3_ROWLEN; e += 16)
				if (base + e < nsamp) cp_async_f(dst + e, src + e);
		}
		cp_async_commit();
	};
	const int ntiles = (p.nsym + K3_TS - 1) / K3_TS;
	if (ntiles > 0) prefetch(0, 0);
	for (int t = 0; t < ntiles; t++) {
		if (t + 1 < ntiles) {
			prefetch((t + 1) & 1, (t + 1) * K3_TS);
			cp_async_wait<1>();
		}
		else cp_async_wait<0>();
		__syncwarp();
		const float2 *my = &mytile[t & 1][half * K3_ROWLEN + phase];
		const int s_end = min(K3_TS, p.nsym - t * K3_TS);
		uint32_t word = 0;
		for (int sl = 0; sl < s_end; sl++) {
			const float2 x = my[sl * 5];
			// (1j)^rot pre-rotation (Demod.cpp:44-65), branch free: swap on odd rot, negate on rot >= 2 (sign flips are exact)
			float re = (rot & 1) ? -x.y : x.x, im = (rot & 1) ? x.x : x.y;
			if (rot & 2) { re = -re; im = -im; }
			rot = (rot + 1) & 3;
			const float tt = __fadd_rn(__fmul_rn(re, cj), __fmul_rn(im, sj));
			hist = (hist << 1) | (tt > 0.0f ? 1u : 0u);
			const float at = fabsf(tt);
			if (p.ps_ema) { // Demod.cpp:67-91
				ma = __fadd_rn(__fmul_rn(weight, ma), __fmul_rn(omw, at));
				const int i0 = (max_idx - 1) & 15;
				const float v0 = __shfl_sync(0xffffffffu, ma, half * 16 + i0);
				const float v1 = __shfl_sync(0xffffffffu, ma, half * 16 + ((i0 + 1) & 15));
				const float v2 = __shfl_sync(0xffffffffu, ma, half * 16 + ((i0 + 2) & 15));
				float mv = v0;
				int best = i0;
				if (v1 > mv) { mv = v1; best = (i0 + 1) & 15; }
				if (v2 > mv) { mv = v2; best = (i0 + 2) & 15; }
				max_idx = best;
			}
			else { // Demod.cpp:129-160: ring slot `last` takes |t|, sums run over slots 0..11 in slot order
#pragma unroll
				for (int l = 0; l < 12; l++) mem[l] = (l == last) ? at : mem[l];
				last = (last + 1) % 12;
				float avg = mem[0];
#pragma unroll
				for (int l = 1; l < 12; l++) avg = __fadd_rn(avg, mem[l]);
				float mv = 0.0f;
				const int prev_max = max_idx;
#pragma unroll
				for (int q = -2; q <= 2; q++) {
					const int jj = (prev_max + q) & 15;
					const float v = __shfl_sync(0xffffffffu, avg, half * 16 + jj);
					if (v > mv) { mv = v; max_idx = jj; }
				}
			}
			const uint32_t hb = __shfl_sync(0xffffffffu, hist, half * 16 + max_idx);
			const uint32_t bit = ((hb >> 3) ^ (hb >> 4)) & 1u; // nDelay = 3 (Model.h:219)
			word |= bit << sl;
			if (p.tap_dec && active && h == 0) p.tap_dec[inst * p.nsym + t * K3_TS + sl] = bit ? 1.0f : -1.0f;
		}
		if (active && h == 0) p.dbits[inst * p.dwords + t] = word;
		if (active && phase == 0 && p.mode_level) { // ScatterPLL level: ((((0+n0)+n1)+n2)+n3)+n4, then / 5
			const float2 *rowt = &mytile[t & 1][half * K3_ROWLEN];
			for (int sl = h; sl < s_end; sl += 16) {
				float acc = 0.0f;
#pragma unroll
				for (int jx = 0; jx < 5; jx++) {
					const float2 x = rowt[sl * 5 + jx];
					acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y)));
				}
				p.lvl[(long long)row * p.lvl_stride + t * K3_TS + sl] = __fdiv_rn(acc, 5.0f);
			}
		}
		__syncwarp();
	}
	if (active) {
		PsState &st = p.ps[inst];
		st.ma[h] = ma;
#pragma unroll
		for (int dd = 0; dd < 5; dd++) {
			const uint32_t pl = __ballot_sync(hmask, (hist >> dd) & 1u) >> (half * 16);
			if (h == 0) st.plane[dd] = pl;
		}
		if (h == 0) { st.max_idx = max_idx; st.rot = rot; st.last = last; }
		if (!p.ps_ema) {
#pragma unroll
			for (int l = 0; l < 12; l++) p.ps_mem[(long long)(h * 12 + l) * ninst + inst] = mem[l];
		}
	}
}

// positions at which Decoder::cannotBeValid (AIS.cpp:111-142) can fire: 30 62 96 168 184 192 336 385 448,
// plus MAX_FRAME_BITS (AIS.h:172) -- one bit per frame position
__constant__ uint32_t c_abort_bits[35];

// Outside a frame the decoder is a tiny automaton; q encodes it in one register:
//   q = 0..5   TRAINING with min(position, 5) alternations seen (only "position > 4" is ever tested, AIS.h:105-113)
//   q = 8..14  STARTFLAG with position = q - 7 (AIS.h:116-137)
__device__ __forceinline__ int dec_q_from_state(const DecState &d) {
	return d.state == ST_TRAINING ? min(d.position, 5) : (d.state == ST_STARTFLAG ? 7 + d.position : 0);
}

// K3b: the five AIS::Decoder instances of one row (AIS.h:91-181) in lanes 0..4 of one warp, one symbol per
// iteration for all of them.  Every lane runs the same straight-line code each symbol -- the out-of-frame automaton
// (one table lookup) and the in-frame bit collector (bits gathered in a register, flushed to shared memory once per
// 32) are both evaluated and masked -- so a row costs the same whether or not it is collecting a frame; only the
// rare events (frame start, word flush, abort positions, closing flag + CRC + Reset vote) branch.
constexpr int DK_THREADS = 128;
template <int MODEL, bool TAPS>
__global__ void __launch_bounds__(DK_THREADS) k_decode(const K3Params p) {
	__shared__ uint32_t frames_all[DK_THREADS / 32][DEC_WORDS * 32];
	__shared__ float tile_all[DK_THREADS / 32][2][K3_ROWLEN]; // MODEL 0: the row's FIR37 samples; MODEL 2: its 32 symbol levels
	__shared__ uint8_t lut_all[DK_THREADS / 32][64];
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const int row = blockIdx.x * (DK_THREADS / 32) + wib;
	if (row >= p.rows) return; // whole warp
	const int phase = lane;
	const bool active = lane < 5;
	float(*tile)[K3_ROWLEN] = tile_all[wib];
	// transition table of the out-of-frame automaton, index (q << 2) | (alt << 1) | Bit:
	// bits 0-3 next q, bit 4 TRAINING->STARTFLAG (start_idx is taken), bit 5 0111111|0 seen: the frame starts
	uint8_t *lut = lut_all[wib];
	for (int i = lane; i < 64; i += 32) {
		const int qq = i >> 2, al = (i >> 1) & 1, bt = i & 1;
		int qn;
		if (qq < 8) qn = al ? min(qq + 1, 5) : (qq == 5 ? 8 + 2 * bt : 0);  // TRAINING (AIS.h:103-114)
		else qn = qq == 14 ? (bt ? 0 : 15) : (bt ? qq + 1 : 0);             // STARTFLAG (AIS.h:116-137)
		const int to_sf = qq < 8 && qn >= 8, enter = qn == 15;
		lut[i] = (uint8_t)((enter ? 0 : qn) | (to_sf << 4) | (enter << 5));
	}
	__syncwarp();

	DecCtx ctx;
	ctx.frame = frames_all[wib] + lane;
	ctx.mode_level = p.mode_level;
	DecState d;
	const int sidx = row * 5 + (active ? phase : 0);
	const long long nthr_total = (long long)p.rows * 5;
	if (active) {
		d = p.dec[sidx];
		for (int w = 0; w < DEC_WORDS; w++) ctx.frame[w * K3_THREADS] = p.dec_data[(long long)w * nthr_total + sidx];
	}
	else {
		d.state = ST_TRAINING; d.lastBit = 0; d.prev = 0; d.position = 0; d.one_seq = 0; d.level = 0.f; d.start_idx = 0;
	}
	const long long clk0 = clock64();
	int n_slow = 0, n_crc = 0, n_crcbits = 0;
	int in_data = active && d.state == ST_DATAFCS;
	int q = dec_q_from_state(d);
	int prev = d.prev, lastBit = d.lastBit;
	int pos = in_data ? d.position : 0, ones = in_data ? d.one_seq : 0;
	float level = d.level;
	uint32_t cur = in_data ? ctx.frame[(pos >> 5) * K3_THREADS] : 0u; // the partially filled frame word
	int start_rel = -1; // slot*5+phase of the most recent TRAINING -> STARTFLAG transition in this submit
	int ntap = 0;
	// slots in which this phase has a sample (Deinterleave forwards partial groups at both ends of a submit)
	const int lo_rel = (int)(p.abs_lo - p.abs_begin), hi_rel = (int)(p.abs_hi - p.abs_begin);
	const int slot_lo = phase >= lo_rel ? 0 : 1;
	const int slot_hi = (hi_rel - phase + 4) / 5;
	const int per_sym = MODEL == 2 ? 1 : 5;
	const int nelem = p.nsym * per_sym;
	const float *src_row = MODEL == 2 ? p.lvl + (long long)row * p.lvl_stride : p.Ef + (long long)row * p.e_stride + p.e_begin;
	auto prefetch = [&](int buf, int s0) {
		const int base = s0 * per_sym;
		for (int e = lane; e < K3_TS * per_sym; e += 32)
			if (base + e < nelem) cp_async_f(&tile[buf][e], src_row + base + e);
		cp_async_commit();
	};
	const int ntiles = (p.nsym + K3_TS - 1) / K3_TS;
	if (ntiles > 0) prefetch(0, 0);
	for (int t = 0; t < ntiles; t++) {
		if (t + 1 < ntiles) {
			prefetch((t + 1) & 1, (t + 1) * K3_TS);
			cp_async_wait<1>();
		}
		else cp_async_wait<0>();
		__syncwarp();
		const int s_end = min(K3_TS, p.nsym - t * K3_TS);
		// decision bits and validity of this lane's 32 slots
		uint32_t dword = 0, vword = 0;
		if (MODEL == 2) {
			if (active) {
				dword = p.dbits[(long long)sidx * p.dwords + t];
				vword = s_end >= 32 ? 0xffffffffu : ((1u << s_end) - 1u);
			}
		}
		else if (active) {
			const float *my = &tile[t & 1][phase];
			for (int sl = 0; sl < s_end; sl++) {
				const int slot = t * K3_TS + sl;
				const float bsmp = my[sl * 5];
				const bool valid = slot >= slot_lo && slot < slot_hi;
				dword |= (bsmp > 0.0f ? 1u : 0u) << sl;
				vword |= (valid ? 1u : 0u) << sl;
				if (TAPS && valid) p.tap_dec[(long long)sidx * p.nsym + ntap++] = bsmp;
			}
		}
		for (int sl = 0; sl < s_end; sl++) {
			const int dd = (dword >> sl) & 1u;
			const int valid = (vword >> sl) & 1u;
			const int Bit = 1 ^ dd ^ prev; // NRZI (AIS.h:93-96)
			const int lastBit_before = lastBit;
			const int alt = Bit ^ lastBit_before;
			const int tv = lut[(q << 2) | (alt << 1) | Bit];
			const int upd = valid & (in_data ^ 1), dat = valid & in_data;
			const int start_before = start_rel;
			const float level_before = level;
			// ---- out of frame: TRAINING / STARTFLAG automaton ----
			const int ev = upd ? (tv >> 4) : 0; // bit 0: start_idx taken, bit 1: the frame starts
			start_rel = (ev & 1) ? (t * K3_TS + sl) * 5 + phase : start_rel;
			q = upd ? (tv & 15) : q;
			// ---- in frame: DATAFCS (AIS.h:141-175) ----
			const int five = ones == 5;
			const int append = dat & ((five & (Bit ^ 1)) ^ 1); // a 0 after five 1s is a stuffing bit and is dropped
			cur |= (uint32_t)(append & Bit) << (pos & 31);
			const int pos_n = pos + append;
			if (MODEL == 2) {
				const float lv = tile[t & 1][sl];
				level = (dat && ctx.mode_level) ? __fadd_rn(level, lv) : level;
			}
			ones = dat ? (Bit ? ones + 1 : 0) : ones;
			const int closing = dat & Bit & five; // sixth 1 in a row: closing flag (AIS.h:151-161)
			const int full = append & ((pos_n & 31) == 0);
			const int abortpos = dat & ((c_abort_bits[pos_n >> 5] >> (pos_n & 31)) & 1u);
			pos = pos_n;
			prev = valid ? dd : prev;
			lastBit = valid ? Bit : lastBit;
			if ((ev >> 1) | full | abortpos) { // lane-local rare events
				if (ev >> 1) { // 0111111|0: the frame starts (AIS.h:120-124)
					in_data = 1;
					q = 0;
					pos = 0; ones = 0; level = 0.0f; cur = 0u;
					d.start_idx = start_rel >= 0 ? p.abs_begin + start_rel : d.start_idx;
					for (int w = 0; w < DEC_WORDS; w++) ctx.frame[w * K3_THREADS] = 0u; // msg.clear()
				}
				if (full) {
					ctx.frame[((pos >> 5) - 1) * K3_THREADS] = cur;
					cur = 0u;
				}
				if (abortpos && !closing) { // position == MaxBits || cannotBeValid(position) (AIS.h:172)
					if (pos & 31) ctx.frame[(pos >> 5) * K3_THREADS] = cur;
					if (pos == MAX_FRAME_BITS || dec_cannot_be_valid(ctx, pos)) { in_data = 0; q = 0; }
				}
			}
			const unsigned closers = __ballot_sync(0xffffffffu, closing);
			if (!closers) continue;
			// ---- some decoder of the row saw a closing flag: CRC, frame emission, Reset of the siblings ----
			n_slow++;
			int fr_len = 0;
			float fr_level = 0.0f;
			bool found = false;
			if (closing) {
				if (pos & 31) ctx.frame[(pos >> 5) * K3_THREADS] = cur;
				fr_level = ctx.mode_level ? __fdiv_rn(level, (float)pos) : 0.0f;
				const int len = pos - 7;
				if (len >= 16 && dec_crc16(ctx, len)) {
					found = true;
					fr_len = len;
				}
				in_data = 0;
				q = 0;
				if (p.dbg) { n_crc++; n_crcbits += len > 0 ? len : 0; }
			}
			const unsigned vote = __ballot_sync(0xffffffffu, found);
			if (vote) { // FOUNDMESSAGE -> Reset to the four sibling decoders (AIS.cpp:47-49,98-108)
				const int winner = __ffs(vote) - 1; // lowest phase runs first (DSP.h:108-112)
				const int rel = (t * K3_TS + sl) * 5 + phase;
				if (lane == winner) {
					float ppm = 0.0f;
					if (MODEL == 2 && p.ppmtab) { // tag.ppm of the CGF block that delivered the group's 5th sample
						const long long last_of_group = p.abs_begin + (long long)(t * K3_TS + sl) * 5 + 4;
						int bi = (int)((last_of_group - p.blk_abs0) >> 9);
						bi = bi < 0 ? 0 : (bi >= p.nblk ? p.nblk - 1 : bi);
						ppm = p.ppmtab[p.stepidx[row * p.nblk + bi]];
					}
					emit_frame(p.ring, p.ring_count, p.ring_cap, p.chunk, p.blk, ctx, row, phase, fr_len, fr_level, ppm, d.start_idx, p.abs_begin + rel);
				}
				else if (active && (lane < winner || !valid)) { // already stepped this symbol (or no sample in this slot), then reset
					in_data = 0;
					q = 0;
				}
				else if (active) { // reset first, then step this symbol from TRAINING/0: only the NRZI memory survives
					in_data = 0;
					level = level_before;
					start_rel = start_before;
					q = alt ? 1 : 0;
				}
			}
		}
		__syncwarp();
	}
	if (p.dbg) {
		const long long dt = clock64() - clk0;
		for (int o = 16; o > 0; o >>= 1) {
			n_crc += __shfl_xor_sync(0xffffffffu, n_crc, o);
			n_crcbits += __shfl_xor_sync(0xffffffffu, n_crcbits, o);
		}
		if (lane == 0) {
			p.dbg[row * 4 + 0] = dt;
			p.dbg[row * 4 + 1] = n_slow;
			p.dbg[row * 4 + 2] = n_crc;
			p.dbg[row * 4 + 3] = n_crcbits;
		}
	}
	if (active) {
		if (in_data) {
			d.state = ST_DATAFCS;
			d.position = pos;
			d.one_seq = ones;
			if (pos & 31) ctx.frame[(pos >> 5) * K3_THREADS] = cur; // keep the partial word with the persisted frame
		}
		else if (q < 8) { d.state = ST_TRAINING; d.position = q; d.one_seq = 0; }
		else { d.state = ST_STARTFLAG; d.position = q - 7; d.one_seq = 0; }
		if (!in_data && q >= 8 && start_rel >= 0) d.start_idx = p.abs_begin + start_rel;
		d.level = level;
		d.prev = prev;
		d.lastBit = lastBit;
		for (int w = 0; w < DEC_WORDS; w++) p.dec_data[(long long)w * nthr_total + sidx] = ctx.frame[w * K3_THREADS];
		p.dec[sidx] = d;
	}
}

// K3b': the same decoders, event driven.  RPW rows per warp (lane = row_in_warp * 5 + phase).  Outside a frame the
// decoder is a regular-expression matcher on the NRZI-decoded bit stream -- ">= 5 alternations, a repeat, the rest of
// 0111 1110" (AIS.h:103-137) -- so a word of 32 symbols is handled with a handful of bitwise operations: the positions
// where TRAINING would enter STARTFLAG are  ~alt & alt<<1 & ... & alt<<5 , and for each of them (about one every other
// word on noise) the outcome of STARTFLAG is read off the following bits.  A failed STARTFLAG at bit f only leaves
// "alternations are counted from f" behind (variable e).  Only when a row has a decoder inside a frame, or a start
// flag completes, or at the first (FM model) / last word of a submit, the row's five decoders drop to the exact
// bit-serial machine of k_decode for that word -- which is where CRC, frame emission and the Reset broadcast
// (AIS.cpp:47-49, Model.cpp:566-573) happen.
constexpr int DK2_WARPS = 2;
template <int MODEL, bool TAPS, int RPW>
__global__ void __launch_bounds__(DK2_WARPS * 32) k_decode2(const K3Params p) {
	constexpr int PER_SYM = 1;
	constexpr int ROWEL = K3_TS; // elements of one row in a tile: the 32 symbol levels (MODEL 2 only)
	__shared__ uint32_t frames_all[DK2_WARPS][DEC_WORDS * 32];
	__shared__ float tile_all[DK2_WARPS][RPW][3][ROWEL];
	__shared__ uint8_t lut_all[DK2_WARPS][64];
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const int g = lane / 5, phase = lane - 5 * g;
	const int row0 = (blockIdx.x * DK2_WARPS + wib) * RPW;
	if (row0 >= p.rows) return; // whole warp
	const int row = row0 + g;
	const bool active = g < RPW && row < p.rows;
	const unsigned gsh = 5 * (g < RPW ? g : 0);
	float(*tile)[3][ROWEL] = tile_all[wib];
	uint8_t *lut = lut_all[wib];
	for (int i = lane; i < 64; i += 32) { // same transition table as k_decode
		const int qq = i >> 2, al = (i >> 1) & 1, bt = i & 1;
		int qn;
		if (qq < 8) qn = al ? min(qq + 1, 5) : (qq == 5 ? 8 + 2 * bt : 0);
		else qn = qq == 14 ? (bt ? 0 : 15) : (bt ? qq + 1 : 0);
		const int to_sf = qq < 8 && qn >= 8, enter = qn == 15;
		lut[i] = (uint8_t)((enter ? 0 : qn) | (to_sf << 4) | (enter << 5));
	}
	__syncwarp();

	DecCtx ctx;
	ctx.frame = frames_all[wib] + lane;
	ctx.mode_level = p.mode_level;
	DecState d;
	const int sidx = active ? row * 5 + phase : 0;
	const long long nthr_total = (long long)p.rows * 5;
	if (active) {
		d = p.dec[sidx];
		for (int w = 0; w < DEC_WORDS; w++) ctx.frame[w * K3_THREADS] = p.dec_data[(long long)w * nthr_total + sidx];
	}
	else {
		d.state = ST_TRAINING; d.lastBit = 0; d.prev = 0; d.position = 0; d.one_seq = 0; d.level = 0.f; d.start_idx = 0;
	}
	int in_data = active && d.state == ST_DATAFCS;
	int q = dec_q_from_state(d);
	int prev = d.prev, lastBit = d.lastBit;
	int pos = in_data ? d.position : 0, ones = in_data ? d.one_seq : 0;
	float level = d.level;
	uint32_t cur = in_data ? ctx.frame[(pos >> 5) * K3_THREADS] : 0u;
	int start_rel = -1;
	int ntap = 0;
	// scan representation of a decoder in TRAINING: alternations are counted for bit indices > e_w (relative to the
	// current word), altprev = alternation flags of the previous word, q_carry = STARTFLAG state a failing flag is in at
	// the word boundary
	bool scanrep = !in_data && q < 8;
	int e_w = -1 - q, q_carry = 0;
	uint32_t altprev = q ? (0xffffffffu << (32 - q)) : 0u;

	const int lo_rel = (int)(p.abs_lo - p.abs_begin), hi_rel = (int)(p.abs_hi - p.abs_begin);
	const int slot_lo = phase >= lo_rel ? 0 : 1;
	const int slot_hi = (hi_rel - phase + 4) / 5;
	const int nelem = p.nsym * PER_SYM;
	auto prefetch = [&](int buf, int s0) {
		const int base = s0 * PER_SYM;
#pragma unroll
		for (int g2 = 0; g2 < RPW; g2++) {
			const int r2 = row0 + g2;
			if (MODEL == 2 && r2 < p.rows) {
				const float *src_row = p.lvl + (long long)r2 * p.lvl_stride;
				for (int e = lane; e < ROWEL; e += 32)
					if (base + e < nelem) cp_async_f(&tile[g2][buf][e], src_row + base + e);
			}
		}
		cp_async_commit();
	};
	const int ntiles = (p.nsym + K3_TS - 1) / K3_TS;
	// decision bits (sample > 0) and slot validity of word t for this lane
	auto get_word = [&](int t, uint32_t &dword, uint32_t &vword) {
		dword = 0;
		vword = 0;
		if (!active) return;
		const int s_end = min(K3_TS, p.nsym - t * K3_TS);
		dword = p.dbits[(long long)sidx * p.dwords + t];
		vword = s_end >= 32 ? 0xffffffffu : ((1u << s_end) - 1u);
		if (MODEL != 2) { // Deinterleave forwards partial groups at both ends of a submit
			const int lo = slot_lo - t * K3_TS, hi = slot_hi - t * K3_TS;
			if (lo > 0) vword &= lo >= 32 ? 0u : (0xffffffffu << lo);
			if (hi < 32) vword &= hi <= 0 ? 0u : ((1u << hi) - 1u);
		}
	};
	uint32_t dword = 0, vword = 0, dnext = 0, vnext = 0;
	uint32_t pre1 = 0, pre2 = 0; // MODEL 2: decision words t+1 and t+2, loaded two iterations before their first use
	auto load_dbits = [&](int t) -> uint32_t { return (active && t < ntiles) ? p.dbits[(long long)sidx * p.dwords + t] : 0u; };
	if (ntiles > 0) {
		prefetch(0, 0);
		if (ntiles > 1) prefetch(1, K3_TS);
		else cp_async_commit();
		cp_async_wait<1>();
		__syncwarp();
		get_word(0, dword, vword);
		if (MODEL == 2) { pre1 = load_dbits(1); pre2 = load_dbits(2); }
	}
	for (int t = 0; t < ntiles; t++) {
		if (t + 2 < ntiles) prefetch((t + 2) % 3, (t + 2) * K3_TS);
		else cp_async_commit();
		cp_async_wait<1>(); // tiles <= t+1 have landed
		__syncwarp();
		dnext = 0;
		vnext = 0;
		if (MODEL == 2) {
			if (active && t + 1 < ntiles) {
				const int se = min(K3_TS, p.nsym - (t + 1) * K3_TS);
				dnext = pre1;
				vnext = se >= 32 ? 0xffffffffu : ((1u << se) - 1u);
			}
			pre1 = pre2;
			pre2 = load_dbits(t + 3);
		}
		else if (t + 1 < ntiles) get_word(t + 1, dnext, vnext);
		(void)ntap;
		const int s_end = min(K3_TS, p.nsym - t * K3_TS);
		const bool last = t == ntiles - 1;
		// ---------------- fast path: all five decoders of the row are in TRAINING ----------------
		bool ser_lane = active && (!scanrep || last || (MODEL != 2 && t == 0));
		bool ser_row = ((__ballot_sync(0xffffffffu, ser_lane) >> gsh) & 31u) != 0;
		uint32_t Bitw = 0, alt = 0;
		int e_new = e_w, qc_new = 0;
		if (active && !ser_row) {
			Bitw = ~(dword ^ ((dword << 1) | (uint32_t)prev));
			alt = Bitw ^ ((Bitw << 1) | (uint32_t)lastBit);
			uint32_t run5 = __funnelshift_l(altprev, alt, 1);
			run5 &= __funnelshift_l(altprev, alt, 2);
			run5 &= __funnelshift_l(altprev, alt, 3);
			run5 &= __funnelshift_l(altprev, alt, 4);
			run5 &= __funnelshift_l(altprev, alt, 5);
			uint32_t E = ~alt & run5; // a repeat after >= 5 alternations: TRAINING -> STARTFLAG if the alternations count
			if (E) {
				const uint32_t Bn = ~(dnext ^ ((dnext << 1) | (dword >> 31)));
				const int avail_next = __popc(vnext); // valid slots are a prefix of the next word
				while (E) {
					const int j = __ffs(E) - 1;
					E &= E - 1;
					if (j - 5 <= e_new) continue; // some of the five alternations precede the last reset
					const int b = (Bitw >> j) & 1;
					const int need = b ? 4 : 6; // ones still to come before the closing 0 of the flag
					unsigned long long U = (((unsigned long long)Bn << 32) | Bitw) >> (j + 1);
					const int avail = 31 - j + avail_next;
					if (avail < 64) U |= ~0ull << avail; // unknown bits must not look like a 0
					const int t1 = ~U ? __ffsll((long long)~U) - 1 : 64; // ones that follow bit j
					const int decide = j + 1 + min(t1, need);   // index of the bit that decides the flag
					if (decide - 32 >= avail_next || t1 == need) { // undecidable here, or 0111 1110 complete: exact machine
						ser_lane = true;
						break;
					}
					e_new = decide; // STARTFLAG failed there: NextState(TRAINING, 0)
					qc_new = decide >= 32 ? 7 + (b ? 3 : 1) + (31 - j) : 0;
				}
			}
		}
		ser_row = ((__ballot_sync(0xffffffffu, ser_lane) >> gsh) & 31u) != 0;
		if (active && !ser_row) { // commit the fast path
			altprev = alt;
			lastBit = (int)(Bitw >> 31);
			prev = (int)(dword >> 31);
			e_w = max(e_new - 32, -64);
			q_carry = qc_new;
		}
		// ---------------- exact path: bit-serial for the rows that need it ----------------
		if (__any_sync(0xffffffffu, active && ser_row)) {
			const bool ser = active && ser_row;
			if (ser && scanrep) { // scan representation -> automaton state at the first bit of the word
				if (q_carry) q = q_carry;
				else {
					const int n_alt = __clz((int)~altprev);
					q = max(0, min(min(5, n_alt), -1 - e_w));
				}
				scanrep = false;
			}
			const uint32_t vw = ser ? vword : 0u;
			for (int sl = 0; sl < s_end; sl++) {
				const int dd = (dword >> sl) & 1u;
				const int valid = (vw >> sl) & 1u;
				const int Bit = 1 ^ dd ^ prev; // NRZI (AIS.h:93-96)
				const int lastBit_before = lastBit;
				const int altb = Bit ^ lastBit_before;
				const int tv = lut[(q << 2) | (altb << 1) | Bit];
				const int upd = valid & (in_data ^ 1), dat = valid & in_data;
				const int start_before = start_rel;
				const float level_before = level;
				const int ev = upd ? (tv >> 4) : 0; // bit 0: start_idx taken, bit 1: the frame starts
				start_rel = (ev & 1) ? (t * K3_TS + sl) * 5 + phase : start_rel;
				q = upd ? (tv & 15) : q;
				const int five = ones == 5;
				const int append = dat & ((five & (Bit ^ 1)) ^ 1); // a 0 after five 1s is a stuffing bit and is dropped
				cur |= (uint32_t)(append & Bit) << (pos & 31);
				const int pos_n = pos + append;
				if (MODEL == 2) {
					const float lv = tile[g < RPW ? g : 0][t % 3][sl];
					level = (dat && ctx.mode_level) ? __fadd_rn(level, lv) : level;
				}
				ones = dat ? (Bit ? ones + 1 : 0) : ones;
				const int closing = dat & Bit & five; // sixth 1 in a row: closing flag (AIS.h:151-161)
				const int full = append & ((pos_n & 31) == 0);
				const int abortpos = dat & ((c_abort_bits[pos_n >> 5] >> (pos_n & 31)) & 1u);
				pos = pos_n;
				prev = valid ? dd : prev;
				lastBit = valid ? Bit : lastBit;
				if ((ev >> 1) | full | abortpos) { // lane-local rare events
					if (ev >> 1) { // 0111111|0: the frame starts (AIS.h:120-124)
						in_data = 1;
						q = 0;
						pos = 0; ones = 0; level = 0.0f; cur = 0u;
						d.start_idx = start_rel >= 0 ? p.abs_begin + start_rel : d.start_idx;
						for (int w = 0; w < DEC_WORDS; w++) ctx.frame[w * K3_THREADS] = 0u; // msg.clear()
					}
					if (full) {
						ctx.frame[((pos >> 5) - 1) * K3_THREADS] = cur;
						cur = 0u;
					}
					if (abortpos && !closing) { // position == MaxBits || cannotBeValid(position) (AIS.h:172)
						if (pos & 31) ctx.frame[(pos >> 5) * K3_THREADS] = cur;
						if (pos == MAX_FRAME_BITS || dec_cannot_be_valid(ctx, pos)) { in_data = 0; q = 0; }
					}
				}
				const unsigned closers = __ballot_sync(0xffffffffu, closing);
				if (!closers) continue;
				// ---- some decoder saw a closing flag: CRC, frame emission, Reset of the siblings ----
				int fr_len = 0;
				float fr_level = 0.0f;
				bool found = false;
				if (closing) {
					if (pos & 31) ctx.frame[(pos >> 5) * K3_THREADS] = cur;
					fr_level = ctx.mode_level ? __fdiv_rn(level, (float)pos) : 0.0f;
					const int len = pos - 7;
					if (len >= 16 && dec_crc16(ctx, len)) {
						found = true;
						fr_len = len;
					}
					in_data = 0;
					q = 0;
				}
				const unsigned vote = (__ballot_sync(0xffffffffu, found) >> gsh) & 31u; // my row's decoders
				if (vote && ser) { // FOUNDMESSAGE -> Reset to the four sibling decoders (AIS.cpp:47-49,98-108)
					const int winner = __ffs(vote) - 1; // lowest phase runs first (DSP.h:108-112)
					const int rel = (t * K3_TS + sl) * 5 + phase;
					if (phase == winner) {
						float ppm = 0.0f;
						if (MODEL == 2 && p.ppmtab) { // tag.ppm of the CGF block that delivered the group's 5th sample
							const long long last_of_group = p.abs_begin + (long long)(t * K3_TS + sl) * 5 + 4;
							int bi = (int)((last_of_group - p.blk_abs0) >> 9);
							bi = bi < 0 ? 0 : (bi >= p.nblk ? p.nblk - 1 : bi);
							ppm = p.ppmtab[p.stepidx[row * p.nblk + bi]];
						}
						emit_frame(p.ring, p.ring_count, p.ring_cap, p.chunk, p.blk, ctx, row, phase, fr_len, fr_level, ppm, d.start_idx, p.abs_begin + rel);
					}
					else if (phase < winner || !valid) { // already stepped this symbol (or no sample in this slot), then reset
						in_data = 0;
						q = 0;
					}
					else { // reset first, then step this symbol from TRAINING/0: only the NRZI memory survives
						in_data = 0;
						level = level_before;
						start_rel = start_before;
						q = altb ? 1 : 0;
					}
				}
			}
			if (ser && !in_data && q < 8) { // back to the scan representation
				scanrep = true;
				e_w = -1 - q;
				altprev = q ? (0xffffffffu << (32 - q)) : 0u;
				q_carry = 0;
			}
		}
		dword = dnext;
		vword = vnext;
		__syncwarp();
	}
	if (active) {
		// the last word of a submit always runs on the exact machine, so q is current here
		if (in_data) {
			d.state = ST_DATAFCS;
			d.position = pos;
			d.one_seq = ones;
			if (pos & 31) ctx.frame[(pos >> 5) * K3_THREADS] = cur;
		}
		else if (q < 8) { d.state = ST_TRAINING; d.position = q; d.one_seq = 0; }
		else { d.state = ST_STARTFLAG; d.position = q - 7; d.one_seq = 0; }
		if (!in_data && q >= 8 && start_rel >= 0) d.start_idx = p.abs_begin + start_rel;
		d.level = level;
		d.prev = prev;
		d.lastBit = lastBit;
		for (int w = 0; w < DEC_WORDS; w++) p.dec_data[(long long)w * nthr_total + sidx] = ctx.frame[w * K3_THREADS];
		p.dec[sidx] = d;
	}
}

// K3c: the five decoders of a row, fully word-parallel.  Every state of AIS::Decoder::Run (AIS.h:91-181) consumes a
// run of bits of the 32-symbol word with bitwise operations instead of one step per bit:
//   TRAINING  : candidate TRAINING->STARTFLAG transitions are  E = ~alt & alt<<1 & .. & alt<<5  (a repeat after five
//               alternations); one counts only if its five alternations come after the last reset (index e).  What
//               STARTFLAG does with it is read off the next bits (count of ones that follow) in the same iteration.
//   STARTFLAG : only when a flag straddles a word boundary: position so far in sfP.
//   DATAFCS   : closing flag = first run of six ones (carry-in `ones` prepended), stuffing bits = zeros after five
//               ones, both by shifted ANDs; the surviving bits are squeezed together and appended to the frame;
//               the cannotBeValid()/MaxBits exits (AIS.cpp:111-142) are evaluated only when the position crosses one
//               of their lengths; the signal level is summed bit by bit in the reference's order.
// A CRC-valid frame is rare; when one closes in a word, the row rolls back to the state at the start of the word
// section, replays it up to that bit (its siblings one bit less if they come later in the round-robin order of
// DSP.h:108-112), applies the Reset broadcast (AIS.cpp:47-49, Model.cpp:566-573) and carries on.
struct Dk3 {
	int mode;        // 0 TRAINING, 1 STARTFLAG, 2 DATAFCS
	int sfP;         // STARTFLAG: position (1..7)
	int pos, ones;   // DATAFCS: position, one_seq_count
	float level;
	uint32_t cur;    // partially filled frame word
	int e;           // TRAINING: alternations count only at bit indices > e (relative to the current word)
	int start_rel;   // slot*5+phase of the latest TRAINING->STARTFLAG transition of this submit, -1 if none
};

__device__ __forceinline__ uint32_t lowmask(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// Consumes bits [i0, i1) of the word.  Returns 32, or the index of the bit at which a CRC-valid frame closed (the
// state is then TRAINING with e = that bit, the frame bits are complete in shared memory, fr_len / fr_level set).
template <bool LEVEL>
__device__ __forceinline__ int dk3_run(Dk3 &st, const DecCtx &ctx, uint32_t Bitw, uint32_t E, int i0, int i1, const float *__restrict__ lvl,
									   int slot0, int phase, int &fr_len, float &fr_level) {
	int i = i0;
	while (i < i1) {
		if (st.mode == 0) {
			uint32_t Em = E & ~lowmask(i) & lowmask(i1);
			bool done = true;
			while (Em) {
				const int j = __ffs(Em) - 1;
				Em &= Em - 1;
				if (j - 5 <= st.e) continue; // some of the five alternations precede the last reset
				// TRAINING -> STARTFLAG at bit j (AIS.h:107-111); position = Bit ? 3 : 1
				st.start_rel = (slot0 + j) * 5 + phase;
				const int b = (Bitw >> j) & 1;
				const int need = b ? 4 : 6; // ones still to come before the 0 that ends the flag
				const int n = i1 - (j + 1);
				const uint32_t W = n > 0 ? ((Bitw >> (j + 1)) | ~lowmask(n)) : 0xffffffffu; // j + 1 may be 32
				const int t1 = ~W ? __ffs(~W) - 1 : 32; // ones that follow
				const int m = min(t1, need);
				if (m >= n) { // the word ends inside the flag
					st.mode = 1;
					st.sfP = (b ? 3 : 1) + n;
					i = i1;
					done = false;
					break;
				}
				const int decide = j + 1 + m;
				if (t1 == need) { // 0111111|0: the frame starts (AIS.h:120-124)
					st.mode = 2;
					st.pos = 0; st.ones = 0; st.level = 0.0f; st.cur = 0u;
					i = decide + 1;
					done = false;
					break;
				}
				st.e = decide; // the flag failed there: NextState(TRAINING, 0)
			}
			if (done) i = i1;
		}
		else if (st.mode == 1) {
			const int n = i1 - i;
			const uint32_t W = (Bitw >> i) | ~lowmask(n);
			const int t1 = ~W ? __ffs(~W) - 1 : 32;
			const int need = 7 - st.sfP;
			const int m = min(t1, need);
			if (m >= n) {
				st.sfP += n;
				i = i1;
			}
			else {
				const int decide = i + m;
				if (t1 == need) {
					st.mode = 2;
					st.pos = 0; st.ones = 0; st.level = 0.0f; st.cur = 0u;
				}
				else {
					st.mode = 0;
					st.e = decide;
				}
				i = decide + 1;
			}
		}
		else {
			const int n = i1 - i;
			const uint32_t W = (Bitw >> i) & lowmask(n);
			const unsigned long long X = ((unsigned long long)W << st.ones) | ((1ull << st.ones) - 1ull); // carried-in ones first
			const unsigned long long R5 = X & (X << 1) & (X << 2) & (X << 3) & (X << 4);
			const unsigned long long R6 = R5 & (X << 5);
			const int c = R6 ? (__ffsll((long long)R6) - 1 - st.ones) : 64; // closing flag: the sixth 1 in a row (AIS.h:151-161)
			const int endb = c < n ? c : n - 1;                              // last bit consumed if no early exit
			const uint32_t Sw = (uint32_t)((~X & (R5 << 1)) >> st.ones) & lowmask(endb + 1); // stuffing zeros
			uint32_t bits = W & lowmask(endb + 1);
			for (uint32_t tmp = Sw; tmp;) { // squeeze the stuffing bits out, highest first
				const int sb = 31 - __clz((int)tmp);
				tmp &= ~(1u << sb);
				bits = (bits & lowmask(sb)) | ((sb >= 31 ? 0u : (bits >> (sb + 1))) << sb);
			}
			const int cnt = endb + 1 - __popc(Sw);
			const int pos0 = st.pos;
			const int sh = pos0 & 31;
			uint32_t cur = st.cur | (bits << sh);
			if (sh + cnt >= 32) {
				ctx.frame[(pos0 >> 5) * K3_THREADS] = cur;
				cur = sh ? (bits >> (32 - sh)) : 0u;
			}
			const int newpos = pos0 + cnt;
			// exits by length: position == MaxBits || cannotBeValid(position), tested after every bit (AIS.h:172)
			int exit_m = -1;
			{
				const int w0 = (pos0 + 1) >> 5, w1 = newpos >> 5;
				bool any = false;
				for (int w = w0; w <= w1 && w < 35; w++) {
					uint32_t ab = c_abort_bits[w];
					if (w == w0) ab &= ~lowmask((pos0 + 1) & 31);
					if (w == w1) ab &= lowmask((newpos & 31) + 1);
					any |= ab != 0;
				}
				if (any) {
					ctx.frame[(newpos >> 5) * K3_THREADS] = cur; // type / mmsi fields must be readable
					for (int Pa = pos0 + 1; Pa <= newpos; Pa++) {
						if (!((c_abort_bits[Pa >> 5] >> (Pa & 31)) & 1u)) continue;
						const int r = Pa - pos0 - 1; // ordinal of the appended bit that makes position == Pa
						int m = r;
						for (;;) {
							const int m2 = r + __popc(Sw & lowmask(m + 1));
							if (m2 == m) break;
							m = m2;
						}
						if (m == c) break; // closing flag on the same bit: NextState(TRAINING) came first
						if (Pa == MAX_FRAME_BITS || dec_cannot_be_valid(ctx, Pa)) {
							exit_m = m;
							break;
						}
					}
				}
			}
			if (exit_m >= 0) {
				st.mode = 0;
				st.e = i + exit_m;
				i += exit_m + 1;
				continue;
			}
			if (LEVEL && ctx.mode_level) { // level += tag.sample_lvl for every bit in DATAFCS, in order (AIS.h:146-147)
				float lv = st.level;
#pragma unroll
				for (int m = 0; m < 32; m++) {
					const float v = lvl[min(i + m, 31)];
					lv = m <= endb ? __fadd_rn(lv, v) : lv;
				}
				st.level = lv;
			}
			if (c < n) { // closing flag
				if (newpos & 31) ctx.frame[(newpos >> 5) * K3_THREADS] = cur;
				st.mode = 0;
				st.e = i + c;
				i += c + 1;
				const int len = newpos - 7;
				if (len >= 16 && dec_crc16_bytes(ctx, len)) {
					fr_len = len;
					fr_level = ctx.mode_level ? __fdiv_rn(st.level, (float)newpos) : 0.0f;
					st.pos = newpos;
					return i - 1;
				}
			}
			else {
				st.pos = newpos;
				st.cur = cur;
				const int tot = n + st.ones; // trailing ones of the consumed bits (a stuffing zero resets the count)
				const unsigned long long Y = ~(X << (64 - tot));
				st.ones = Y ? __clzll((long long)Y) : tot;
				i = i1;
			}
		}
	}
	return 32;
}

constexpr int DK3_WARPS = 2;
template <int MODEL, int RPW>
__global__ void __launch_bounds__(DK3_WARPS * 32) k_decode3(const K3Params p) {
	__shared__ uint32_t frames_all[DK3_WARPS][DEC_WORDS * 32];
	__shared__ float tile_all[DK3_WARPS][RPW][3][K3_TS];
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const int g = lane / 5, phase = lane - 5 * g;
	const int row0 = (blockIdx.x * DK3_WARPS + wib) * RPW;
	if (row0 >= p.rows) return; // whole warp
	const int row = row0 + g;
	const bool active = g < RPW && row < p.rows;
	const int gbase = 5 * (g < RPW ? g : 0);
	float(*tile)[3][K3_TS] = tile_all[wib];

	DecCtx ctx;
	ctx.frame = frames_all[wib] + lane;
	ctx.mode_level = p.mode_level;
	DecState d;
	const int sidx = active ? row * 5 + phase : 0;
	const long long nthr_total = (long long)p.rows * 5;
	Dk3 st;
	st.mode = 0; st.sfP = 0; st.pos = 0; st.ones = 0; st.level = 0.0f; st.cur = 0u; st.e = -1; st.start_rel = -1;
	int prev = 0, lastBit = 0;
	uint32_t altprev = 0u;
	if (active) {
		d = p.dec[sidx];
		prev = d.prev;
		lastBit = d.lastBit;
		if (d.state == ST_DATAFCS) {
			st.mode = 2;
			st.pos = d.position;
			st.ones = d.one_seq;
			st.level = d.level;
			const int nw = (d.position >> 5) + 1;
			for (int w = 0; w < nw && w < DEC_WORDS; w++) ctx.frame[w * K3_THREADS] = p.dec_data[(long long)w * nthr_total + sidx];
			st.cur = (d.position & 31) ? ctx.frame[(d.position >> 5) * K3_THREADS] : 0u;
		}
		else if (d.state == ST_STARTFLAG) {
			st.mode = 1;
			st.sfP = d.position;
		}
		else { // TRAINING with `position` alternations counted so far (only "> 4" is ever tested)
			const int q = min(d.position, 5);
			st.e = -1 - q;
			altprev = q ? (0xffffffffu << (32 - q)) : 0u;
		}
	}
	const int lo_rel = (int)(p.abs_lo - p.abs_begin), hi_rel = (int)(p.abs_hi - p.abs_begin);
	const int slot_lo = phase >= lo_rel ? 0 : 1; // Deinterleave forwards partial groups at both ends of a submit
	const int slot_hi = (hi_rel - phase + 4) / 5;
	const int ntiles = (p.nsym + K3_TS - 1) / K3_TS;
	auto prefetch = [&](int buf, int s0) {
		if (MODEL == 2) {
#pragma unroll
			for (int g2 = 0; g2 < RPW; g2++) {
				const int r2 = row0 + g2;
				if (r2 < p.rows && s0 + lane < p.nsym) cp_async_f(&tile[g2][buf][lane], p.lvl + (long long)r2 * p.lvl_stride + s0 + lane);
			}
		}
		cp_async_commit();
	};
	auto load_dbits = [&](int t) -> uint32_t { return (active && t < ntiles) ? p.dbits[(long long)sidx * p.dwords + t] : 0u; };
	uint32_t pre0 = load_dbits(0), pre1 = load_dbits(1), pre2 = load_dbits(2);
	if (ntiles > 0) {
		prefetch(0, 0);
		if (ntiles > 1) prefetch(1, K3_TS);
		else cp_async_commit();
	}
	int nbits_total = 0; // valid bits seen by this lane in this submit
	for (int t = 0; t < ntiles; t++) {
		if (t + 2 < ntiles) prefetch((t + 2) % 3, (t + 2) * K3_TS);
		else cp_async_commit();
		cp_async_wait<2>(); // tile t has landed
		__syncwarp();
		uint32_t dword = pre0;
		pre0 = pre1;
		pre1 = pre2;
		pre2 = load_dbits(t + 3);
		// valid slots of this word for this lane: [lo, hi)
		int lo = max(0, slot_lo - t * K3_TS), hi = min(K3_TS, min(p.nsym, slot_hi) - t * K3_TS);
		if (!active) hi = 0;
		const int nb = max(0, hi - lo);
		const int slot0 = t * K3_TS + lo;
		dword >>= lo;
		const uint32_t Bitw = ~(dword ^ ((dword << 1) | (uint32_t)prev)); // NRZI (AIS.h:93-96)
		const uint32_t alt = Bitw ^ ((Bitw << 1) | (uint32_t)lastBit);
		uint32_t run5 = __funnelshift_l(altprev, alt, 1);
		run5 &= __funnelshift_l(altprev, alt, 2);
		run5 &= __funnelshift_l(altprev, alt, 3);
		run5 &= __funnelshift_l(altprev, alt, 4);
		run5 &= __funnelshift_l(altprev, alt, 5);
		const uint32_t E = ~alt & run5 & lowmask(nb);
		const float *lvl = &tile[g < RPW ? g : 0][t % 3][lo];
		int i = 0;
		for (;;) {
			const Dk3 saved = st;
			const int i_saved = i;
			int fr_len = 0;
			float fr_level = 0.0f;
			const int x = dk3_run<MODEL == 2>(st, ctx, Bitw, E, i, nb, lvl, slot0, phase, fr_len, fr_level);
			i = x < 32 ? x + 1 : nb;
			if (!__any_sync(0xffffffffu, x < 32)) break;
			// a frame with a good CRC closed somewhere in the warp: per row, the first one in (bit, phase) order wins
			const int key = x < 32 ? (x + lo) * 8 + phase : 0x7fffffff; // bit index in slot units (lanes of a row may differ in lo)
			int rowmin = 0x7fffffff;
#pragma unroll
			for (int k2 = 0; k2 < 5; k2++) rowmin = min(rowmin, __shfl_sync(0xffffffffu, key, gbase + k2));
			if (rowmin == 0x7fffffff || !active) continue; // nothing in this row: its lanes have finished the word already
			const int xs = rowmin >> 3, pw = rowmin & 7; // slot (relative to the word) and phase of the winner
			if (key == rowmin) { // FOUNDMESSAGE: publish, Reset goes to the four siblings (AIS.cpp:47-49,98-108)
				float ppm = 0.0f;
				const int slot = t * K3_TS + xs;
				if (MODEL == 2 && p.ppmtab) { // tag.ppm of the CGF block that delivered the group's 5th sample
					const long long last_of_group = p.abs_begin + (long long)slot * 5 + 4;
					int bi = (int)((last_of_group - p.blk_abs0) >> 9);
					bi = bi < 0 ? 0 : (bi >= p.nblk ? p.nblk - 1 : bi);
					ppm = p.ppmtab[p.stepidx[row * p.nblk + bi]];
				}
				const long long sidx0 = st.start_rel >= 0 ? p.abs_begin + st.start_rel : d.start_idx;
				const int slotw = atomicAdd(p.ring_count, 1);
				if (slotw < p.ring_cap) {
					FrameRec &r = p.ring[slotw];
					r.row = row; r.phase = phase; r.nbits = fr_len - 16; r.level = fr_level; r.ppm = ppm; r.chunk = p.chunk; r.blk = p.blk;
					r.start_idx = sidx0;
					r.end_idx = p.abs_begin + (long long)slot * 5 + phase;
					const int nw = (st.pos + 31) >> 5;
					for (int w = 0; w < DEC_WORDS; w++) r.data[w] = w < nw ? frame_word(ctx, w) : 0u; // msg.clear() left the rest zero
				}
			}
			else { // sibling: replay up to the winner's bit, then Reset -> NextState(TRAINING, 0)
				st = saved;
				const int xl = xs - lo; // the winner's slot as a bit index of this lane's word (may be -1 when lo = 1)
				const int stop = max(i_saved, min(nb, phase < pw ? xl + 1 : xl)); // earlier phases have already stepped that symbol
				int fl = 0;
				float fv = 0.0f;
				if (stop > i_saved) dk3_run<MODEL == 2>(st, ctx, Bitw, E, i_saved, stop, lvl, slot0, phase, fl, fv);
				st.mode = 0;
				st.e = stop - 1;
				i = stop;
			}
		}
		if (nb > 0) {
			altprev = nb >= 32 ? alt : ((alt << (32 - nb)) | (altprev >> nb)); // keep "bit 31 = latest alternation flag"
			lastBit = (int)((Bitw >> (nb - 1)) & 1u);
			prev = (int)((dword >> (nb - 1)) & 1u);
			st.e = max(st.e - nb, -64);
			nbits_total += nb;
		}
		__syncwarp();
	}
	if (active) {
		if (st.mode == 2) {
			d.state = ST_DATAFCS;
			d.position = st.pos;
			d.one_seq = st.ones;
			if (st.pos & 31) ctx.frame[(st.pos >> 5) * K3_THREADS] = st.cur;
			const int nw = (st.pos >> 5) + 1;
			for (int w = 0; w < nw && w < DEC_WORDS; w++) p.dec_data[(long long)w * nthr_total + sidx] = ctx.frame[w * K3_THREADS];
		}
		else if (st.mode == 1) { d.state = ST_STARTFLAG; d.position = st.sfP; d.one_seq = 0; }
		else { // TRAINING: alternations counted = trailing alternation flags that come after the last reset
			const int n_alt = __clz((int)~altprev);
			d.state = ST_TRAINING;
			d.position = max(0, min(min(5, n_alt), -1 - st.e));
			d.one_seq = 0;
		}
		if (st.mode != 0 && st.start_rel >= 0) d.start_idx = p.abs_begin + st.start_rel;
		d.level = st.level;
		d.prev = prev;
		d.lastBit = lastBit;
		p.dec[sidx] = d;
	}
}

// ModelBase: SimplePLL (DSP.cpp:28-57) + one Decoder per row; strictly sequential per row.
struct PllState { int prev; float pll; int fast; };
__global__ void __launch_bounds__(K3_THREADS) k_base(const float *__restrict__ Ef, long long e_stride, int e_begin, int n, int rows,
													   PllState *__restrict__ pll, DecState *__restrict__ dec, uint32_t *__restrict__ dec_data, FrameRec *__restrict__ ring,
													   int *__restrict__ ring_count, int ring_cap, int chunk, int blk, float *__restrict__ tap_dec, int *__restrict__ tap_cnt) {
	__shared__ uint32_t frames[DEC_WORDS * K3_THREADS];
	const int tid = threadIdx.x;
	const int row = blockIdx.x * K3_THREADS + tid;
	if (row >= rows) return;
	DecCtx ctx;
	ctx.frame = frames + tid;
	ctx.mode_level = 1;
	DecState d = dec[row * 5];
	const long long nthr_total = (long long)rows * 5;
	for (int w = 0; w < DEC_WORDS; w++) frames[w * K3_THREADS + tid] = dec_data[(long long)w * nthr_total + row * 5];
	PllState pl = pll[row];
	const float *e = Ef + (long long)row * e_stride + e_begin;
	int ntap = 0;
	for (int i = 0; i < n; i++) {
		const float x = e[i];
		const int bit = x > 0.0f;
		if (bit != pl.prev) pl.pll = __fadd_rn(pl.pll, __fmul_rn(__fsub_rn(0.5f, pl.pll), pl.fast ? 0.6f : 0.05f));
		pl.pll = __fadd_rn(pl.pll, 0.2f);
		if (pl.pll >= 1.0f) {
			if (tap_dec) tap_dec[(long long)row * n + ntap++] = x;
			int fr_len = 0, lb = 0;
			float fr_level = 0.f;
			const bool found = dec_step(d, ctx, x, 0.0f, 0, fr_len, fr_level, lb);
			if (found) emit_frame(ring, ring_count, ring_cap, chunk, blk, ctx, row, 0, fr_len, fr_level, 0.0f, d.start_idx, 0);
			// DecoderMessage -> SimplePLL::Signal (Model.cpp:434-435; DSP.cpp:46-57): the last NextState decides
			pl.fast = (d.state == ST_TRAINING) ? 1 : (d.state == ST_STARTFLAG ? 0 : pl.fast);
			pl.pll = __fsub_rn(pl.pll, (float)(int)pl.pll);
		}
		pl.prev = bit;
	}
	for (int w = 0; w < DEC_WORDS; w++) dec_data[(long long)w * nthr_total + row * 5] = frames[w * K3_THREADS + tid];
	dec[row * 5] = d;
	pll[row] = pl;
	if (tap_cnt) tap_cnt[row] = ntap;
}

} // namespace aisgpu
