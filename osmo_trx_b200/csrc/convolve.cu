// convolve.cu — batched convolve_real / convolve_complex / base_convolve_* for sm_100a.
//
// Same contract as arch/common/convolve.h:4-26: y[i] = sum_k x[i + start - (h_len-1) + k] * h[k].
// These replace arch/x86/convolve_sse_3.c (SSE3) and arch/arm/convolve_neon.S: one thread per
// output sample, taps staged in shared memory, burst rows read through the read-only path.  The
// summation trees of the SSE3 kernels are reproduced so
// decision-bearing callers get bit-identical results; `base` selects the sequential MAC of
// arch/common/convolve_base.c:27-82.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {
namespace {

// real taps: lane-ordered sum for one component (c = 0 re, 1 im); xx points at the first tap's sample
__device__ __forceinline__ float real_tree(const float2 *__restrict__ xx, const float *hs, int h_len, int c)
{
	float L[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
#define PX(k) fm((c ? __ldg(&xx[(k)]).y : __ldg(&xx[(k)]).x), hs[2 * (k)])
		switch (h_len) {
		case 4: L[j] = PX(j); break;
		case 8: L[j] = fa(PX(j), PX(4 + j)); break;
		case 12: L[j] = fa(fa(PX(j), PX(4 + j)), PX(8 + j)); break;
		case 16: L[j] = fa(fa(PX(j), PX(4 + j)), fa(PX(8 + j), PX(12 + j))); break;
		case 20: L[j] = fa(fa(fa(PX(j), PX(4 + j)), PX(8 + j)), fa(PX(12 + j), PX(16 + j))); break;
		default: {
			float a = 0.0f;
			for (int g = 0; g < h_len / 4; g++)
				a = fa(a, PX(4 * g + j));
			L[j] = a;
		}
		}
#undef PX
	}
	return fa(fa(L[0], L[1]), fa(L[2], L[3]));
}

} // namespace

// mode: 0 real SSE order, 1 complex SSE order, 2 real sequential, 3 complex sequential
__global__ void __launch_bounds__(256)
convolve_kernel(const float *__restrict__ x, int x_stride, const float *__restrict__ h, int h_len,
		float *__restrict__ y, int y_stride, int start, int len, int n, int mode)
{
	extern __shared__ float hs[]; // interleaved complex taps
	for (int t = threadIdx.x; t < 2 * h_len; t += blockDim.x)
		hs[t] = h[t];
	__syncthreads();
	const long total = (long)n * len;
	for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
		const int b = (int)(o / len), i = (int)(o % len);
		const float2 *xx = reinterpret_cast<const float2 *>(x) + (size_t)b * x_stride + (i + start - (h_len - 1));
		float2 r = make_float2(0.0f, 0.0f);
		const bool seq = (mode >= 2) || (h_len % 4);
		if (seq) {
			if (mode == 0 || mode == 2) {
				for (int k = 0; k < h_len; k++) {
					const float2 v = __ldg(&xx[k]);
					r.x = fa(r.x, fm(v.x, hs[2 * k]));
					r.y = fa(r.y, fm(v.y, hs[2 * k]));
				}
			} else {
				for (int k = 0; k < h_len; k++) {
					const float2 v = __ldg(&xx[k]);
					r.x = fa(r.x, fs(fm(v.x, hs[2 * k]), fm(v.y, hs[2 * k + 1])));
					r.y = fa(r.y, fa(fm(v.x, hs[2 * k + 1]), fm(v.y, hs[2 * k])));
				}
			}
		} else if (mode == 0) {
			r.x = real_tree(xx, hs, h_len, 0);
			r.y = real_tree(xx, hs, h_len, 1);
		} else {
			float Lr[4], Li[4];
			const bool two = !(h_len % 8);
#pragma unroll
			for (int j = 0; j < 4; j++) {
				float ar = 0.0f, ai = 0.0f, br = 0.0f, bi = 0.0f;
				if (two) {
					for (int g = 0; g < h_len; g += 8) {
						float2 v = __ldg(&xx[g + j]);
						float hr = hs[2 * (g + j)], hi = hs[2 * (g + j) + 1];
						ar = fa(ar, fs(fm(hr, v.x), fm(hi, v.y)));
						ai = fa(ai, fa(fm(hr, v.y), fm(hi, v.x)));
						v = __ldg(&xx[g + 4 + j]);
						hr = hs[2 * (g + 4 + j)]; hi = hs[2 * (g + 4 + j) + 1];
						br = fa(br, fs(fm(hr, v.x), fm(hi, v.y)));
						bi = fa(bi, fa(fm(hr, v.y), fm(hi, v.x)));
					}
					Lr[j] = fa(ar, br);
					Li[j] = fa(ai, bi);
				} else {
					for (int g = 0; g < h_len; g += 4) {
						const float2 v = __ldg(&xx[g + j]);
						const float hr = hs[2 * (g + j)], hi = hs[2 * (g + j) + 1];
						ar = fa(ar, fs(fm(hr, v.x), fm(hi, v.y)));
						ai = fa(ai, fa(fm(hr, v.y), fm(hi, v.x)));
					}
					Lr[j] = ar;
					Li[j] = ai;
				}
			}
			r.x = fa(fa(Lr[0], Lr[1]), fa(Lr[2], Lr[3]));
			r.y = fa(fa(Li[0], Li[1]), fa(Li[2], Li[3]));
		}
		reinterpret_cast<float2 *>(y)[(size_t)b * y_stride + i] = r;
	}
}

} // namespace trxb200
