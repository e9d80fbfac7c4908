// filterbank.cu — polyphase Resampler (Resampler.cpp:131-150) and M-channel Channelizer / Synthesis
// filterbanks (Channelizer.cpp:74-99, Synthesis.cpp:85-114) for sm_100a.
//
// Resampler: out[i] = sum_k in[(q*i)/p - (L-1) + k] * part[(q*i)%p][k]; the reference issues one
// convolve_real(len=1) per output, here every output is a thread and the per-path taps stay in L1.
//
// Channelizer: the reference deinterleaves the wideband block into M branches (commutator reversed),
// runs a 16-tap real FIR per branch and an M-point forward DFT across branches for every time index.
// The kernel tiles time: a CTA stages the FIR outputs of all M branches for T time indices in shared
// memory (the 16-sample history comes from the previous call's tail, kept per object), then lanes = time,
// warps = output channel evaluate the DFT with twiddles broadcast from shared memory, so the wideband
// input is read once and every channel row is written coalesced.  Blocks of one call are consecutive in
// time, so n_blocks can be processed per launch with bit-identical FIR results.
// Synthesis is the mirror: DFT across channels first (recomputed for the 15-sample halo from the kept
// input tail), then the branch FIRs, written interleaved.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {
namespace {

// sse_conv_real* summation order for one output; x points at tap 0's sample, taps are real
__device__ __forceinline__ float2 fir_real_exact(const float2 *__restrict__ x, const float *__restrict__ h, int L)
{
	float2 r;
	if (L % 4) {
		float ar = 0.0f, ai = 0.0f;
		for (int k = 0; k < L; k++) {
			const float2 v = x[k];
			ar = fa(ar, fm(v.x, h[k]));
			ai = fa(ai, fm(v.y, h[k]));
		}
		return make_float2(ar, ai);
	}
	float Lr[4], Li[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
#define PR(k) fm(x[(k)].x, h[(k)])
#define PI(k) fm(x[(k)].y, h[(k)])
		switch (L) {
		case 4: Lr[j] = PR(j); Li[j] = PI(j); break;
		case 8: Lr[j] = fa(PR(j), PR(4 + j)); Li[j] = fa(PI(j), PI(4 + j)); break;
		case 12: Lr[j] = fa(fa(PR(j), PR(4 + j)), PR(8 + j)); Li[j] = fa(fa(PI(j), PI(4 + j)), PI(8 + j)); break;
		case 16:
			Lr[j] = fa(fa(PR(j), PR(4 + j)), fa(PR(8 + j), PR(12 + j)));
			Li[j] = fa(fa(PI(j), PI(4 + j)), fa(PI(8 + j), PI(12 + j)));
			break;
		case 20:
			Lr[j] = fa(fa(fa(PR(j), PR(4 + j)), PR(8 + j)), fa(PR(12 + j), PR(16 + j)));
			Li[j] = fa(fa(fa(PI(j), PI(4 + j)), PI(8 + j)), fa(PI(12 + j), PI(16 + j)));
			break;
		default: {
			float ar = 0.0f, ai = 0.0f;
			for (int g = 0; g < L / 4; g++) { ar = fa(ar, PR(4 * g + j)); ai = fa(ai, PI(4 * g + j)); }
			Lr[j] = ar; Li[j] = ai;
		}
		}
#undef PR
#undef PI
	}
	r.x = fa(fa(Lr[0], Lr[1]), fa(Lr[2], Lr[3]));
	r.y = fa(fa(Li[0], Li[1]), fa(Li[2], Li[3]));
	return r;
}

} // namespace

__global__ void __launch_bounds__(256)
resampler_kernel(const float *__restrict__ in, int in_stride, float *__restrict__ out, int out_len, int out_stride,
		 int n_streams, int p, int q, int L, const float *__restrict__ parts)
{
	const long total = (long)n_streams * out_len;
	for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
		const int s = (int)(o / out_len), i = (int)(o % out_len);
		const int n = (int)(((long)q * i) / p), path = (int)(((long)q * i) % p);
		const float2 *x = reinterpret_cast<const float2 *>(in) + (size_t)s * in_stride + (n - (L - 1));
		reinterpret_cast<float2 *>(out)[(size_t)s * out_stride + i] = fir_real_exact(x, parts + (size_t)path * L, L);
	}
}

// ---- channelizer ----
// in: [n_blocks*block_len][m] wideband samples (time-major), hist_in: [m][L] previous tail per branch,
// out: [m][n_blocks*block_len].  grid.x tiles time by T = 32.
__global__ void __launch_bounds__(256)
channelizer_kernel(const float *__restrict__ in, const float *__restrict__ hist_in, float *__restrict__ out,
		   int m, int L, long total_t, const float *__restrict__ sub, const float2 *__restrict__ tw)
{
	extern __shared__ __align__(16) float2 csm[];
	float2 *y = csm;			// [m][33]
	float2 *w = csm + (size_t)m * 33;	// [m] twiddles
	const int T = 32;
	for (int k = threadIdx.x; k < m; k += blockDim.x) w[k] = tw[k];
	const float2 *xin = reinterpret_cast<const float2 *>(in);
	const float2 *hin = reinterpret_cast<const float2 *>(hist_in);
	for (long t0 = (long)blockIdx.x * T; t0 < total_t; t0 += (long)gridDim.x * T) {
		__syncthreads();
		// branch FIRs: work item = (branch r, time t0+tt); lanes walk branches so the wideband read is coalesced
		for (int it = threadIdx.x; it < m * T; it += blockDim.x) {
			const int r = it % m, tt = it / m;
			const long t = t0 + tt;
			float2 acc = make_float2(0.0f, 0.0f);
			if (t < total_t) {
				float2 xs[32];
				const int col = m - 1 - r; // deinterleave: in[i*m + n] -> branch m-1-n (Channelizer.cpp:44-45)
				for (int k = 0; k < L; k++) {
					const long ti = t - (L - 1) + k;
					xs[k] = (ti >= 0) ? __ldg(&xin[ti * m + col]) : hin[(size_t)r * L + (L + ti)];
				}
				acc = fir_real_exact(xs, sub + (size_t)r * L, L);
			}
			y[r * 33 + tt] = acc;
		}
		__syncthreads();
		// DFT across branches: lanes = time, warps stride over channels
		const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
		const long t = t0 + lane;
		for (int c = wid; c < m; c += nw) {
			float ar = 0.0f, ai = 0.0f;
			int idx = 0;
			for (int r = 0; r < m; r++) {
				const float2 v = y[r * 33 + lane], ww = w[idx];
				ar = fmaf(v.x, ww.x, ar); ar = fmaf(-v.y, ww.y, ar);
				ai = fmaf(v.x, ww.y, ai); ai = fmaf(v.y, ww.x, ai);
				idx += c; if (idx >= m) idx -= m;
			}
			if (t < total_t)
				reinterpret_cast<float2 *>(out)[(size_t)c * total_t + t] = make_float2(ar, ai);
		}
	}
}

// new history = last L samples of every branch (Channelizer.cpp:88); total_t >= L is guaranteed by block_len >= L
__global__ void channelizer_hist_kernel(const float *__restrict__ in, float *__restrict__ hist_out, int m, int L, long total_t)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m * L) return;
	const int r = i / L, k = i % L;
	const long t = total_t - L + k;
	reinterpret_cast<float2 *>(hist_out)[(size_t)r * L + k] = reinterpret_cast<const float2 *>(in)[t * m + (m - 1 - r)];
}

// ---- synthesis ----
// in: [m][total_t] per-channel samples, tail_in: [m][L] previous input tail per channel, out: [total_t][m]
__global__ void __launch_bounds__(256)
synthesis_kernel(const float *__restrict__ in, const float *__restrict__ tail_in, float *__restrict__ out, int m, int L,
		 long total_t, const float *__restrict__ sub, const float2 *__restrict__ tw)
{
	extern __shared__ __align__(16) float2 ssm[];
	const int T = 32, TW = T + 32; // columns: L-1 halo + T
	float2 *v = ssm;			 // [m][TW+1] DFT outputs per branch over time
	float2 *w = ssm + (size_t)m * (TW + 1);
	float2 *xs = w + m;			 // [m][TW+1] staged inputs
	for (int k = threadIdx.x; k < m; k += blockDim.x) w[k] = tw[k];
	const float2 *xin = reinterpret_cast<const float2 *>(in);
	const float2 *tin = reinterpret_cast<const float2 *>(tail_in);
	const int halo = L - 1;
	for (long t0 = (long)blockIdx.x * T; t0 < total_t; t0 += (long)gridDim.x * T) {
		__syncthreads();
		const int ncol = halo + T;
		for (int it = threadIdx.x; it < m * ncol; it += blockDim.x) {
			const int c = it / ncol, j = it % ncol;
			const long t = t0 - halo + j;
			float2 val = make_float2(0.0f, 0.0f);
			if (t >= 0) { if (t < total_t) val = __ldg(&xin[(size_t)c * total_t + t]); }
			else val = tin[(size_t)c * L + (L + t)];
			xs[c * (TW + 1) + j] = val;
		}
		__syncthreads();
		// forward DFT across channels for every staged column: lanes = column, warps stride over branches
		const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
		for (int j0 = 0; j0 < ncol; j0 += 32) {
			const int j = j0 + lane;
			for (int r = wid; r < m; r += nw) {
				float ar = 0.0f, ai = 0.0f;
				int idx = 0;
				if (j < ncol) {
					for (int c = 0; c < m; c++) {
						const float2 x = xs[c * (TW + 1) + j], ww = w[idx];
						ar = fmaf(x.x, ww.x, ar); ar = fmaf(-x.y, ww.y, ar);
						ai = fmaf(x.x, ww.y, ai); ai = fmaf(x.y, ww.x, ai);
						idx += r; if (idx >= m) idx -= m;
					}
					v[r * (TW + 1) + j] = make_float2(ar, ai);
				}
			}
		}
		__syncthreads();
		// branch FIRs, written interleaved out[t*m + r] (Synthesis.cpp:38-49): lanes walk branches
		for (int it = threadIdx.x; it < m * T; it += blockDim.x) {
			const int r = it % m, tt = it / m;
			const long t = t0 + tt;
			if (t < total_t) {
				const float2 y = fir_real_exact(&v[r * (TW + 1) + tt], sub + (size_t)r * L, L);
				reinterpret_cast<float2 *>(out)[t * m + r] = y;
			}
		}
	}
}

__global__ void synthesis_tail_kernel(const float *__restrict__ in, float *__restrict__ tail_out, int m, int L, long total_t)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m * L) return;
	const int c = i / L, k = i % L;
	reinterpret_cast<float2 *>(tail_out)[(size_t)c * L + k] = reinterpret_cast<const float2 *>(in)[(size_t)c * total_t + (total_t - L + k)];
}

} // namespace trxb200
