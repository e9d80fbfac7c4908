// misc.cu — small batched helpers of sigProcLib.h (energyDetect, vectorSlicer, delayVector) and the
// int16 <-> float converters of arch/common/convert.h.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

// energyDetect (sigProcLib.cpp:1573-1585): sequential float sum of |x[4i]|^2, i < window; one thread per burst
__global__ void energy_detect_kernel(const float *__restrict__ bursts, int stride, int blen, int n, unsigned window,
				     float *__restrict__ energy)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n) return;
	if (window == 0) { energy[b] = 0.0f; return; }
	if (window > (unsigned)blen) window = blen;
	const float2 *x = reinterpret_cast<const float2 *>(bursts) + (size_t)b * stride;
	float e = 0.0f;
	for (unsigned i = 0; i < window; i++)
		e = fa(e, norm2(__ldg(&x[4 * i])));
	energy[b] = e / (float)window;
}

// vectorSlicer (sigProcLib.cpp:546-556): 0.5*(s+1) computed in double, clamped to [0,1].  The double product of
// 0.5 and a float is exact and rounds back to the same float as the single-precision product (no flush to zero in
// this build), and the comparisons against 1.0 / 0.0 see the same value either way, so the float form is bit-identical.
__device__ __forceinline__ float slice1(float s)
{
	float v = fm(0.5f, fa(s, 1.0f));
	if (v > 1.0f) v = 1.0f;
	else if (v < 0.0f) v = 0.0f;
	return v;
}

// vec: both pointers 16-byte aligned -> one float4 per work item, the (len & 3) tail by block 0
__global__ void __launch_bounds__(256)
vector_slicer_kernel(float *dst, const float *src, size_t len, int vec)
{
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	if (!vec) {
		for (size_t i = tid; i < len; i += nth) dst[i] = slice1(src[i]);
		return;
	}
	const size_t nv = len >> 2;
	const float4 *s4 = reinterpret_cast<const float4 *>(src);
	float4 *d4 = reinterpret_cast<float4 *>(dst);
	size_t i = tid;
	for (; i + nth < nv; i += 2 * nth) {
		float4 a = __ldcs(&s4[i]), b = __ldcs(&s4[i + nth]);
		a.x = slice1(a.x); a.y = slice1(a.y); a.z = slice1(a.z); a.w = slice1(a.w);
		b.x = slice1(b.x); b.y = slice1(b.y); b.z = slice1(b.z); b.w = slice1(b.w);
		__stcs(&d4[i], a);
		__stcs(&d4[i + nth], b);
	}
	if (i < nv) {
		float4 a = __ldcs(&s4[i]);
		a.x = slice1(a.x); a.y = slice1(a.y); a.z = slice1(a.z); a.w = slice1(a.w);
		__stcs(&d4[i], a);
	}
	if (blockIdx.x == 0 && threadIdx.x < (len & 3)) {
		const size_t k = (nv << 2) + threadIdx.x;
		dst[k] = slice1(src[k]);
	}
}

// delayVector (sigProcLib.cpp:1046-1098), exact two-stage form: 20-tap fractional filter in
// sse_conv_real20 order (NO_DELAY span: zero padding both sides), then integer shift with zero fill.
__global__ void __launch_bounds__(256)
delay_vector_kernel(const float *__restrict__ in, int stride, int len, int n, const float *__restrict__ delay,
		    float *__restrict__ out, int out_stride, float delay_scale = 1.0f)
{
	for (int b = blockIdx.x; b < n; b += gridDim.x) {
		const float dly = fm(delay[b], delay_scale); // (1.0 or -1.0 / -sps: exact)
		const int whole = (int)floorf(dly);
		const float frac = fs(dly, (float)whole);
		const bool use_f = (double)fabsf(frac) > 1e-2;
		int f = 0;
		if (use_f) f = min(max((int)floorf(fm(frac, 64.0f)), 0), 63);
		const float2 *x = reinterpret_cast<const float2 *>(in) + (size_t)b * stride;
		float2 *o = reinterpret_cast<float2 *>(out) + (size_t)b * out_stride;
		for (int i = threadIdx.x; i < len; i += blockDim.x) {
			const int m = i - whole; // shifted[i] = y[i - whole]
			float2 r = make_float2(0.0f, 0.0f);
			if (m >= 0 && m < len) {
				if (!use_f) {
					r = __ldg(&x[m]);
				} else {
					float pr[20], pi[20];
#pragma unroll
					for (int k = 0; k < 20; k++) {
						const int idx = m - 9 + k;
						float2 v = make_float2(0.0f, 0.0f);
						if (idx >= 0 && idx < len) v = __ldg(&x[idx]);
						pr[k] = fm(v.x, c_tab.delay[f][k]);
						pi[k] = fm(v.y, c_tab.delay[f][k]);
					}
					float Lr[4], Li[4];
#pragma unroll
					for (int j = 0; j < 4; j++) {
						Lr[j] = fa(fa(fa(pr[j], pr[4 + j]), pr[8 + j]), fa(pr[12 + j], pr[16 + j]));
						Li[j] = fa(fa(fa(pi[j], pi[4 + j]), pi[8 + j]), fa(pi[12 + j], pi[16 + j]));
					}
					r = make_float2(fa(fa(Lr[0], Lr[1]), fa(Lr[2], Lr[3])), fa(fa(Li[0], Li[1]), fa(Li[2], Li[3])));
				}
			}
			o[i] = r;
		}
	}
}

// delay_vector_blk_kernel — the same arithmetic, register blocked like convolve_blk_kernel (convolve.cu): a warp owns
// 128 consecutive outputs of one burst, stages the 147 input samples they depend on (zero outside the vector) into a
// transposed shared window, and each lane evaluates 4 consecutive outputs from 23 samples in registers with the
// burst's 20 taps read warp-uniformly from __constant__.  The integer shift only moves the window.
__global__ void __launch_bounds__(256)
delay_vector_blk_kernel(const float *__restrict__ in, int stride, int len, int n, const float *__restrict__ delay,
			float *__restrict__ out, int out_stride, float negzero, float delay_scale = 1.0f)
{
	constexpr int NW = kCvTile + 19, KL = (NW + 31) / 32;
	__shared__ __align__(16) float2 win_all[8][4 * kCvPitch];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	float2 *win = win_all[warp];
	const float2 NZ = bc2(negzero);
	const int tpr = (len + kCvTile - 1) / kCvTile;
	const int ntiles = n * tpr;
	float2 nxt[KL];
	int whole_n = 0, f_n = -1;
	auto fetch = [&](int t_) {
		const int b_ = t_ / tpr, i0_ = (t_ - b_ * tpr) * kCvTile;
		const float dly = fm(delay[b_], delay_scale); // (1.0 or -1.0 / -sps: exact)
		whole_n = (int)floorf(dly);
		const float frac = fs(dly, (float)whole_n);
		f_n = -1;
		if ((double)fabsf(frac) > 1e-2) f_n = min(max((int)floorf(fm(frac, 64.0f)), 0), 63);
		const float2 *x = reinterpret_cast<const float2 *>(in) + (size_t)b_ * stride;
		const int x0 = i0_ - whole_n - 9; // input index of window sample 0
#pragma unroll
		for (int k = 0; k < KL; k++) {
			const int xi = x0 + lane + 32 * k;
			nxt[k] = make_float2(0.0f, 0.0f);
			if (lane + 32 * k < NW && xi >= 0 && xi < len) nxt[k] = __ldg(&x[xi]);
		}
	};
	const int tstep = gridDim.x * 8;
	int t = blockIdx.x * 8 + warp;
	if (t < ntiles) fetch(t);
	for (; t < ntiles; t += tstep) {
		const int b = t / tpr, i0 = (t - b * tpr) * kCvTile;
		const int whole = whole_n, f = f_n;
		__syncwarp();
#pragma unroll
		for (int k = 0; k < KL; k++) {
			const int sidx = lane + 32 * k;
			if (sidx < NW) win[(sidx & 3) * kCvPitch + (sidx >> 2)] = nxt[k];
		}
		__syncwarp();
		if (t + tstep < ntiles) fetch(t + tstep);
		float2 w[23];
#pragma unroll
		for (int j = 0; j < 23; j++) w[j] = win[(j & 3) * kCvPitch + lane + (j >> 2)];
		float2 res[4];
		if (f >= 0) {
			const float *hd = c_tab.delay[f];
#pragma unroll
			for (int r = 0; r < 4; r++) {
				float2 L[4];
#pragma unroll
				for (int j = 0; j < 4; j++) {
#define PX(k) mul2(w[r + (k)], bc2(hd[(k)]), NZ)
					L[j] = add2(add2(add2(PX(j), PX(4 + j)), PX(8 + j)), add2(PX(12 + j), PX(16 + j)));
#undef PX
				}
				res[r] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
			}
		} else {
#pragma unroll
			for (int r = 0; r < 4; r++) res[r] = w[r + 9];
		}
		float2 *o = reinterpret_cast<float2 *>(out) + (size_t)b * out_stride + i0 + 4 * lane;
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const int m = i0 + 4 * lane + r - whole; // shifted[i] = y[i - whole], zero where that falls outside the vector
			if (m < 0 || m >= len) res[r] = make_float2(0.0f, 0.0f);
		}
		const int left = len - (i0 + 4 * lane);
		if (left >= 4 && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
			reinterpret_cast<float4 *>(o)[0] = make_float4(res[0].x, res[0].y, res[1].x, res[1].y);
			reinterpret_cast<float4 *>(o)[1] = make_float4(res[2].x, res[2].y, res[3].x, res[3].y);
		} else {
#pragma unroll
			for (int r = 0; r < 4; r++)
				if (r < left) o[r] = res[r];
		}
	}
}

// convert_float_short (arch/x86/convert_sse_3.c / convert_sse_4_1.c semantics: multiply, convert with
// round-to-nearest-even, saturate to int16)
__device__ __forceinline__ int f2s1(float x, float scale)
{
	const float v = fm(x, scale);
	if (!(v == v) || v >= 2147483648.0f || v < -2147483648.0f) return -32768;
	return max(-32768, min(32767, __float2int_rn(v)));
}
__device__ __forceinline__ unsigned f2s2(float a, float b, float scale)
{
	return ((unsigned)f2s1(a, scale) & 0xffffu) | ((unsigned)f2s1(b, scale) << 16);
}

// base_convert_float_short (arch/common/convert_base.c:20-25): `short = float * scale` as gcc compiles it for x86-64 -
// cvttss2si (truncation; 0x80000000 for NaN and values outside int32), low 16 bits kept
__device__ __forceinline__ int f2s1_trunc(float x, float scale)
{
	const float v = fm(x, scale);
	if (!(v == v) || v >= 2147483648.0f || v < -2147483648.0f) return 0;
	return (int)(short)(__float2int_rz(v) & 0xffff);
}

// vec: both pointers 16-byte aligned -> eight values per work item (two 16-byte loads, one 16-byte store)
// mode 0: SSE semantics for every element; 1: the x86 dispatcher of convert_float_short (arch/x86/convert.c:63-71,
// convert_sse_3.c:38-47): SSE for whole groups of eight, the scalar truncating loop for the len % 8 tail; 2: scalar everywhere
__global__ void __launch_bounds__(256)
convert_float_short_kernel(int16_t *__restrict__ out, const float *__restrict__ in, float scale, size_t len, int vec, int mode)
{
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	if (mode) {
		const size_t body = mode == 1 ? (len & ~(size_t)7) : 0;
		for (size_t i = tid; i < len; i += nth) out[i] = (int16_t)(i < body ? f2s1(in[i], scale) : f2s1_trunc(in[i], scale));
		return;
	}
	if (!vec) {
		for (size_t i = tid; i < len; i += nth) out[i] = (int16_t)f2s1(in[i], scale);
		return;
	}
	const size_t nv = len >> 3;
	const float4 *s4 = reinterpret_cast<const float4 *>(in);
	uint4 *d4 = reinterpret_cast<uint4 *>(out);
	for (size_t i = tid; i < nv; i += nth) {
		const float4 a = __ldcs(&s4[2 * i]), b = __ldcs(&s4[2 * i + 1]);
		uint4 r;
		r.x = f2s2(a.x, a.y, scale); r.y = f2s2(a.z, a.w, scale);
		r.z = f2s2(b.x, b.y, scale); r.w = f2s2(b.z, b.w, scale);
		__stcs(&d4[i], r);
	}
	if (blockIdx.x == 0 && threadIdx.x < (len & 7)) {
		const size_t k = (nv << 3) + threadIdx.x;
		out[k] = (int16_t)f2s1(in[k], scale);
	}
}

__device__ __forceinline__ float2 s2f2(unsigned w) { return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16)); }

__global__ void __launch_bounds__(256)
convert_short_float_kernel(float *__restrict__ out, const int16_t *__restrict__ in, size_t len, int vec)
{
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	if (!vec) {
		for (size_t i = tid; i < len; i += nth) out[i] = (float)in[i];
		return;
	}
	const size_t nv = len >> 3;
	const uint4 *s4 = reinterpret_cast<const uint4 *>(in);
	float4 *d4 = reinterpret_cast<float4 *>(out);
	for (size_t i = tid; i < nv; i += nth) {
		const uint4 a = __ldcs(&s4[i]);
		const float2 p0 = s2f2(a.x), p1 = s2f2(a.y), p2 = s2f2(a.z), p3 = s2f2(a.w);
		__stcs(&d4[2 * i], make_float4(p0.x, p0.y, p1.x, p1.y));
		__stcs(&d4[2 * i + 1], make_float4(p2.x, p2.y, p3.x, p3.y));
	}
	if (blockIdx.x == 0 && threadIdx.x < (len & 7)) {
		const size_t k = (nv << 3) + threadIdx.x;
		out[k] = (float)in[k];
	}
}

} // namespace trxb200
