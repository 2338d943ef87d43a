// fe_common.cuh -- sample-format conversion shared by the front-end kernels (Utilities/Convert.cpp:255-286).
#pragma once
#include "exact.cuh"

namespace aisgpu {

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

} // namespace aisgpu
