// The common-case path of fd_atan2f_common (ais-catcher_b200/csrc/exact.cuh) restated in C and checked against the C library, bit for bit
// (tests/test_exact_host.py compiles and runs it with -ffp-contract=off).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
// returns 1 and *out when the fast path applies
static int atan2_fast(float y, float x, float *out) {
	const uint32_t hx = f2u(x), hy = f2u(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
	const int k = ((int32_t)iy - (int32_t)ix) >> 23;
	if (!((ix - 0x00800000u) < 0x7f000000u && (iy - 0x00800000u) < 0x7f000000u && (uint32_t)(k + 28) <= 51u && hx != 0x3f800000u)) return 0;
	const float t = fabsf(y / x);
	const uint32_t it = f2u(t);
	const int small = it < 0x3ee00000u, c0 = it < 0x3f300000u, c1 = it < 0x3f980000u, c2 = it < 0x401c0000u;
	const float a = c0 ? 2.0f * t : t;
	const float nm = a - (c1 ? 1.0f : 1.5f);
	const float b = (c2 && !c1) ? 1.5f * t : t;
	const float dn = b + (c0 ? 2.0f : 1.0f);
	const float num = small ? t : (c2 ? nm : -1.0f);
	const float den = small ? 1.0f : (c2 ? dn : t);
	const float hi = small ? 0.0f : (c0 ? 4.6364760399e-01f : (c1 ? 7.8539812565e-01f : (c2 ? 9.8279368877e-01f : 1.5707962513e+00f)));
	const float lo = small ? 0.0f : (c0 ? 5.0121582440e-09f : (c1 ? 3.7748947079e-08f : (c2 ? 3.4473217170e-08f : 7.5497894159e-08f)));
	const float xr = num / den;
	static const float aT[11] = { 3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
		-7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f };
	const float z = xr * xr, w = z * z;
	const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
	const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
	const float xs = xr * (s1 + s2);
	const float zz = hi - ((xs - lo) - xr); // small: 0 - ((xs - 0) - xr) == xr - xs
	const float pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
	float r;
	if ((int32_t)hx >= 0) r = u2f(f2u(zz) | (hy & 0x80000000u));
	else {
		const float v = (zz - pi_lo) - pi;                     // the y < 0 result; pi - (z - pi_lo) is its exact negation
		r = u2f(f2u(v) ^ (~hy & 0x80000000u));
	}
	*out = r;
	return 1;
}
int main(int argc, char **argv) {
	long n = argc > 1 ? atol(argv[1]) : 100000000L, fast = 0, bad = 0;
	uint64_t s = 88172645463325252ULL;
	for (long i = 0; i < n; i++) {
		s ^= s << 13; s ^= s >> 7; s ^= s << 17;
		uint32_t a = (uint32_t)s, b = (uint32_t)(s >> 32);
		float y, x;
		switch (i & 3) {
		case 0: y = u2f(a); x = u2f(b); break;                                 // any bit patterns
		case 1: y = u2f((a & 0x807fffffu) | ((100 + (a >> 23) % 56) << 23)); x = u2f((b & 0x807fffffu) | ((100 + (b >> 23) % 56) << 23)); break; // moderate exponents
		case 2: y = u2f((a & 0x807fffffu) | (127u << 23)); x = u2f((b & 0x807fffffu) | ((120 + (b >> 23) % 14) << 23)); break;  // ratios around the reduction thresholds
		default: y = (float)((int32_t)a) * 1e-9f; x = (float)((int32_t)b) * 1e-9f; break;
		}
		float r;
		if (!atan2_fast(y, x, &r)) continue;
		fast++;
		const float want = atan2f(y, x);
		if (f2u(r) != f2u(want)) { if (bad < 10) printf("MISMATCH y=%a x=%a got=%a want=%a\n", y, x, r, want); bad++; }
	}
	printf("checked %ld of %ld on the fast path, mismatches %ld\n", fast, n, bad);
	return bad != 0;
}
