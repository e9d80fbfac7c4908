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

// vectorSlicer (sigProcLib.cpp:546-556): 0.5*(s+1) computed in double, clamped to [0,1]
__global__ void vector_slicer_kernel(float *__restrict__ dst, const float *__restrict__ src, size_t len)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
		float v = (float)(0.5 * (double)fa(src[i], 1.0f));
		if ((double)v > 1.0) v = 1.0f;
		else if ((double)v < 0.0) v = 0.0f;
		dst[i] = v;
	}
}

// delayVector (sigProcLib.cpp:1046-1098), exact two-stage form: 20-tap fractional filter in
// sse_conv_real20 order (NO_DELAY span: zero padding both sides), then integer shift with zero fill.
__global__ void __launch_bounds__(256)
delay_vector_kernel(const float *__restrict__ in, int stride, int len, int n, const float *__restrict__ delay,
		    float *__restrict__ out, int out_stride)
{
	for (int b = blockIdx.x; b < n; b += gridDim.x) {
		const float dly = delay[b];
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

// convert_float_short (arch/x86/convert_sse_3.c / convert_sse_4_1.c semantics: multiply, convert with
// round-to-nearest-even, saturate to int16)
__global__ void convert_float_short_kernel(int16_t *__restrict__ out, const float *__restrict__ in, float scale, size_t len)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
		const float v = fm(in[i], scale);
		int r;
		if (!(v == v) || v >= 2147483648.0f || v < -2147483648.0f) r = -32768;
		else r = max(-32768, min(32767, __float2int_rn(v)));
		out[i] = (int16_t)r;
	}
}

__global__ void convert_short_float_kernel(float *__restrict__ out, const int16_t *__restrict__ in, size_t len)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x)
		out[i] = (float)in[i];
}

} // namespace trxb200
