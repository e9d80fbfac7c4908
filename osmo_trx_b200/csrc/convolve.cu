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

// ---------------------------------------------------------------------------------------------
// convolve_blk_kernel — the SSE-order cases with h_len in {4, 8, .., 24}, register blocked: a warp owns a tile of
// 128 consecutive outputs of one row, stages the 128 + h_len - 1 samples it needs with coalesced loads into a
// transposed shared window (sample s in plane s & 3 at index s >> 2: the per-lane reads below are conflict free),
// and every lane evaluates 4 consecutive outputs from h_len + 3 samples held in registers.  Each output is the same
// expression tree as above on the packed FP32 pipe ((re, im) pairs; exact, see mul2 in detect.cu), so the bits do
// not change; per output the kernel issues 1/3 of the instructions and 1/6 of the load wavefronts of convolve_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kCvTile = 128;  // outputs per warp tile
constexpr int kCvPitch = 52;  // plane pitch in samples (>= (128 + 23 + 3) / 4, = 4 mod 16: conflict-free staging stores)

template <int HLEN, bool CPLX>
__global__ void __launch_bounds__(256)
convolve_blk_kernel(const float *__restrict__ x, int x_stride, const float *__restrict__ h, float *__restrict__ y, int y_stride,
		    int start, int len, int n, float negzero)
{
	__shared__ __align__(16) float2 win_all[8][4 * kCvPitch];
	__shared__ __align__(16) float2 hs[HLEN];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x < HLEN) hs[threadIdx.x] = reinterpret_cast<const float2 *>(h)[threadIdx.x];
	__syncthreads();
	float2 *win = win_all[warp];
	const float2 NZ = bc2(negzero);
	const int tpr = (len + kCvTile - 1) / kCvTile; // tiles per row
	const int ntiles = n * tpr; // the launcher keeps n * tpr inside 31 bits (32-bit index arithmetic below)
	constexpr int NW = kCvTile + HLEN - 1; // samples a tile needs
	constexpr int KL = (NW + 31) / 32;
	// the samples of the NEXT tile are fetched into registers while the current tile is evaluated
	float2 nxt[KL];
	auto fetch = [&](int t_) {
		const int b_ = t_ / tpr, i0_ = (t_ - b_ * tpr) * kCvTile;
		const float2 *xr = reinterpret_cast<const float2 *>(x) + (size_t)b_ * x_stride + (i0_ + start - (HLEN - 1));
		const int nvalid = min(NW, len - i0_ + HLEN - 1); // samples of the tile inside the range the contract covers
#pragma unroll
		for (int k = 0; k < KL; k++) {
			const int sidx = lane + 32 * k;
			nxt[k] = make_float2(0.0f, 0.0f);
			if (sidx < nvalid) nxt[k] = __ldg(&xr[sidx]);
		}
	};
	const int tstep = gridDim.x * 8;
	int t = blockIdx.x * 8 + warp;
	if (t < ntiles) fetch(t);
	for (; t < ntiles; t += tstep) {
		const int b = t / tpr, i0 = (t - b * tpr) * kCvTile;
		__syncwarp();
#pragma unroll
		for (int k = 0; k < KL; k++) {
			const int sidx = lane + 32 * k;
			if (sidx < NW) win[(sidx & 3) * kCvPitch + (sidx >> 2)] = nxt[k];
		}
		__syncwarp();
		if (t + tstep < ntiles) fetch(t + tstep);
		float2 w[HLEN + 3];
#pragma unroll
		for (int j = 0; j < HLEN + 3; j++) w[j] = win[(j & 3) * kCvPitch + lane + (j >> 2)];
		float2 out[4];
#pragma unroll
		for (int r = 0; r < 4; r++) {
			float2 L[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if constexpr (!CPLX) {
#define PX(k) mul2(w[r + (k)], bc2(hs[(k)].x), NZ)
					if constexpr (HLEN == 4) L[j] = PX(j);
					else if constexpr (HLEN == 8) L[j] = add2(PX(j), PX(4 + j));
					else if constexpr (HLEN == 12) L[j] = add2(add2(PX(j), PX(4 + j)), PX(8 + j));
					else if constexpr (HLEN == 16) L[j] = add2(add2(PX(j), PX(4 + j)), add2(PX(8 + j), PX(12 + j)));
					else if constexpr (HLEN == 20) L[j] = add2(add2(add2(PX(j), PX(4 + j)), PX(8 + j)), add2(PX(12 + j), PX(16 + j)));
					else {
						float2 a = make_float2(0.0f, 0.0f);
#pragma unroll
						for (int g = 0; g < HLEN / 4; g++) a = add2(a, PX(4 * g + j));
						L[j] = a;
					}
#undef PX
				} else {
#define TAP(k) cmul_tap(w[r + (k)], bc2(hs[(k)].x), make_float2(hs[(k)].y, -hs[(k)].y), NZ)
					float2 a = make_float2(0.0f, 0.0f);
					if constexpr (HLEN % 8 == 0) {
						float2 c = make_float2(0.0f, 0.0f);
#pragma unroll
						for (int g = 0; g < HLEN; g += 8) {
							a = add2(a, TAP(g + j));
							c = add2(c, TAP(g + 4 + j));
						}
						L[j] = add2(a, c);
					} else {
#pragma unroll
						for (int g = 0; g < HLEN; g += 4) a = add2(a, TAP(g + j));
						L[j] = a;
					}
#undef TAP
				}
			}
			out[r] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
		}
		float2 *yr = reinterpret_cast<float2 *>(y) + (size_t)b * y_stride + i0 + 4 * lane;
		const int left = len - (i0 + 4 * lane);
		if (left >= 4 && (reinterpret_cast<uintptr_t>(yr) & 15u) == 0) {
			reinterpret_cast<float4 *>(yr)[0] = make_float4(out[0].x, out[0].y, out[1].x, out[1].y);
			reinterpret_cast<float4 *>(yr)[1] = make_float4(out[2].x, out[2].y, out[3].x, out[3].y);
		} else {
#pragma unroll
			for (int r = 0; r < 4; r++)
				if (r < left) yr[r] = out[r];
		}
	}
}

template <bool CPLX>
static bool launch_convolve_blk(int h_len, int grid, cudaStream_t st, const float *x, int x_stride, const float *h, float *y,
				int y_stride, int start, int len, int n)
{
	switch (h_len) {
#define CV_CASE(H) case H: convolve_blk_kernel<H, CPLX><<<grid, 256, 0, st>>>(x, x_stride, h, y, y_stride, start, len, n, -0.0f); return true;
	CV_CASE(4) CV_CASE(8) CV_CASE(12) CV_CASE(16) CV_CASE(20) CV_CASE(24)
#undef CV_CASE
	default: return false;
	}
}

} // namespace trxb200
