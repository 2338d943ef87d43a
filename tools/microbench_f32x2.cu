// Microbenchmark: throughput and bit-exactness of packed add.rn.f32x2 (SASS FADD2) vs scalar FADD on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/microbench_f32x2 tools/microbench_f32x2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
	unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rc;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
	return *reinterpret_cast<float2 *>(&rc);
}

template <int PACKED>
__global__ void k_thr(float2 *out, int iters, float2 seed) {
	float2 a[8];
#pragma unroll
	for (int i = 0; i < 8; i++) a[i] = make_float2(seed.x * (i + 1) + threadIdx.x, seed.y * (i + 2));
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (PACKED) a[i] = add2(a[i], a[(i + 1) & 7]);
			else a[i] = make_float2(__fadd_rn(a[i].x, a[(i + 1) & 7].x), __fadd_rn(a[i].y, a[(i + 1) & 7].y));
		}
	}
	float2 s = a[0];
#pragma unroll
	for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_eq(const float2 *a, const float2 *b, int n, unsigned long long *bad) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float2 p = add2(a[i], b[i]);
	float2 s = make_float2(__fadd_rn(a[i].x, b[i].x), __fadd_rn(a[i].y, b[i].y));
	if (__float_as_uint(p.x) != __float_as_uint(s.x) || __float_as_uint(p.y) != __float_as_uint(s.y)) atomicAdd(bad, 1ULL);
}

int main() {
	cudaDeviceProp pr;
	cudaGetDeviceProperties(&pr, 0);
	const int blocks = pr.multiProcessorCount * 8, threads = 256, iters = 4096;
	float2 *out;
	cudaMalloc(&out, sizeof(float2) * blocks * threads);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int packed = 0; packed < 2; packed++) {
		for (int rep = 0; rep < 3; rep++) {
			cudaEventRecord(e0);
			if (packed) k_thr<1><<<blocks, threads>>>(out, iters, make_float2(1e-3f, 2e-3f));
			else k_thr<0><<<blocks, threads>>>(out, iters, make_float2(1e-3f, 2e-3f));
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms;
			cudaEventElapsedTime(&ms, e0, e1);
			double lane_adds = 2.0 * 8 * (double)iters * blocks * threads;
			if (rep == 2)
				printf("%s: %.3f ms, %.2f T lane-adds/s, %.1f lane-adds/clk/SM @%d MHz nominal\n", packed ? "FADD2 (add.rn.f32x2)" : "FADD  (scalar)   ", ms,
					   lane_adds / ms / 1e9, lane_adds / (ms * 1e-3) / pr.multiProcessorCount / (pr.clockRate * 1e3), pr.clockRate / 1000);
		}
	}
	// bit-exactness on raw bit patterns (denormals, infs, nans excluded from compare only when both NaN)
	const int n = 1 << 24;
	std::vector<unsigned> ha(2 * n), hb(2 * n);
	unsigned long long s = 88172645463325252ULL;
	for (int i = 0; i < 2 * n; i++) {
		s ^= s << 13; s ^= s >> 7; s ^= s << 17;
		ha[i] = (unsigned)s;
		hb[i] = (unsigned)(s >> 32);
		if ((i & 3) == 0) { ha[i] &= 0x807fffffu; }           // denormal operand
		if ((i & 7) == 1) { hb[i] = ha[i] ^ 0x80000001u; }    // near-cancellation
		if (((ha[i] >> 23) & 0xff) == 0xff) ha[i] &= ~0x00800000u; // avoid NaN payload differences
		if (((hb[i] >> 23) & 0xff) == 0xff) hb[i] &= ~0x00800000u;
	}
	float2 *da, *db;
	unsigned long long *dbad, hbad = 0;
	cudaMalloc(&da, 8ull * n);
	cudaMalloc(&db, 8ull * n);
	cudaMalloc(&dbad, 8);
	cudaMemcpy(da, ha.data(), 8ull * n, cudaMemcpyHostToDevice);
	cudaMemcpy(db, hb.data(), 8ull * n, cudaMemcpyHostToDevice);
	cudaMemset(dbad, 0, 8);
	k_eq<<<(n + 255) / 256, 256>>>(da, db, n, dbad);
	cudaMemcpy(&hbad, dbad, 8, cudaMemcpyDeviceToHost);
	printf("bit-exactness add.rn.f32x2 vs __fadd_rn on %d pairs: %llu mismatches\n", n, hbad);
	return hbad != 0;
}
