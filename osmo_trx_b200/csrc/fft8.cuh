// fft8.cuh — 8-point forward DFT in registers (radix-2 decimation in frequency), the building block of the
// 64-point branch transform of the Channelizer / Synthesis filterbanks (M = 8 x 8 Cooley-Tukey split).
// Replaces the FFTW plan of ChannelizerBase::initFFT (ChannelizerBase.cpp:141-167; arch/common/fft.c:55-114);
// FFTW is unpinned in the reference, the parity bar is the mathematical DFT at 1e-4 relative.
// Plain float arithmetic on the host (unit-tested there), packed f32x2 add/sub on sm_100.
#pragma once
#include <cuda_runtime.h>

namespace trxb200 {

__host__ __device__ __forceinline__ float2 c_add(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
	unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
	return *reinterpret_cast<float2 *>(&rd);
#else
	return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 c_sub(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
	unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
	return *reinterpret_cast<float2 *>(&rd);
#else
	return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * b (complex), contraction allowed
__host__ __device__ __forceinline__ float2 c_mul(float2 a, float2 b)
{
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * (-j)
__host__ __device__ __forceinline__ float2 c_mul_mj(float2 a) { return make_float2(a.y, -a.x); }

// 4-point forward DFT of (c0, c1, c2, c3) -> (y0, y1, y2, y3)
__host__ __device__ __forceinline__ void fft4(float2 &c0, float2 &c1, float2 &c2, float2 &c3)
{
	const float2 s0 = c_add(c0, c2), s1 = c_sub(c0, c2), s2 = c_add(c1, c3), s3 = c_mul_mj(c_sub(c1, c3));
	c0 = c_add(s0, s2);
	c2 = c_sub(s0, s2);
	c1 = c_add(s1, s3);
	c3 = c_sub(s1, s3);
}

// X[k] = sum_n a[n] e^{-2 pi j n k / 8}, natural order in and out
__host__ __device__ __forceinline__ void fft8(float2 (&a)[8])
{
	const float h = 0.70710678118654752f;
	float2 e0 = c_add(a[0], a[4]), e1 = c_add(a[1], a[5]), e2 = c_add(a[2], a[6]), e3 = c_add(a[3], a[7]);
	float2 o0 = c_sub(a[0], a[4]), d1 = c_sub(a[1], a[5]), d2 = c_sub(a[2], a[6]), d3 = c_sub(a[3], a[7]);
	// odd branch twiddles W8^n: n = 1: (1-j)/sqrt2, n = 2: -j, n = 3: (-1-j)/sqrt2
	float2 o1 = make_float2((d1.x + d1.y) * h, (d1.y - d1.x) * h);
	float2 o2 = c_mul_mj(d2);
	float2 o3 = make_float2((d3.y - d3.x) * h, -(d3.x + d3.y) * h);
	fft4(e0, e1, e2, e3); // X[0], X[2], X[4], X[6]
	fft4(o0, o1, o2, o3); // X[1], X[3], X[5], X[7]
	a[0] = e0; a[2] = e1; a[4] = e2; a[6] = e3;
	a[1] = o0; a[3] = o1; a[5] = o2; a[7] = o3;
}

} // namespace trxb200
