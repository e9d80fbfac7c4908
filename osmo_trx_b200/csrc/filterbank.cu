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
#include "fft8.cuh"

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

// 16-tap fast path: a CTA stages the input span of P polyphase periods of one stream in shared memory;
// thread = (path, period group) keeps its path's 16 taps in registers for the whole kernel, walks the periods
// of the tile and evaluates sse_conv_real16's tree on the packed FP32 pipe (exact: no contraction), so the
// result stays bit-identical to Resampler::rotate.  Lanes walk consecutive outputs: coalesced stores, and the
// 16 window reads of a half-warp fall into one bank sweep as long as q < 2p (decimating ratios keep the
// one-thread-per-output kernel above, whose strided reads the L1 serves better than a bank-conflicted row).
__global__ void __launch_bounds__(256)
resampler16_kernel(const float *__restrict__ in, int in_stride, float *__restrict__ out, int out_len, int out_stride,
		   int n_streams, int p, int q, int P, const float *__restrict__ parts, float negzero)
{
	extern __shared__ __align__(16) float2 rsm[];
	const int tid = threadIdx.x;
	const int G = 256 / p, rho = tid % p, grp = tid / p;
	const bool active = grp < G;
	const float2 nz = make_float2(negzero, negzero);
	float h[16];
	const int path = (int)(((long)q * rho) % p), nrel = (int)(((long)q * rho) / p);
#pragma unroll
	for (int k = 0; k < 16; k++) h[k] = __ldg(&parts[(size_t)path * 16 + k]);
	const int periods_total = out_len / p;
	const int tiles_per_stream = (periods_total + P - 1) / P;
	const long total_tiles = (long)n_streams * tiles_per_stream;
	// two window buffers: the next tile's input span is copied (cp.async, 8 bytes per sample) while this one is evaluated
	const int bufslots = P * q + 15;
	const unsigned rsm_s = (unsigned)__cvta_generic_to_shared(rsm);
	auto issue = [&](long tile_, int buf_) {
		const int s_ = (int)(tile_ / tiles_per_stream), per0_ = (int)(tile_ % tiles_per_stream) * P;
		const int np_ = min(P, periods_total - per0_);
		const float2 *src = reinterpret_cast<const float2 *>(in) + (size_t)s_ * in_stride + ((long)per0_ * q - 15);
		const int cnt = np_ * q + 15;
		const unsigned dst = rsm_s + 8u * (unsigned)(buf_ * bufslots);
		for (int idx = tid; idx < cnt; idx += 256)
			asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * (unsigned)idx), "l"(src + idx) : "memory");
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	int buf = 0;
	if ((long)blockIdx.x < total_tiles) issue(blockIdx.x, 0);
	for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, buf ^= 1) {
		const int s = (int)(tile / tiles_per_stream), per0 = (int)(tile % tiles_per_stream) * P;
		const int np = min(P, periods_total - per0);
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads(); // this tile's span has landed; everybody is done with the other buffer
		if (tile + gridDim.x < total_tiles) issue(tile + gridDim.x, buf ^ 1);
		const float2 *rs = rsm + (size_t)buf * bufslots;
		if (!active) continue;
		float2 *orow = reinterpret_cast<float2 *>(out) + (size_t)s * out_stride + (size_t)per0 * p + rho;
		for (int per = grp; per < np; per += G) {
			const int nb = per * q + nrel;
			float2 L[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const float2 p0 = mul2(rs[nb + j], bc2(h[j]), nz);
				const float2 p1 = mul2(rs[nb + 4 + j], bc2(h[4 + j]), nz);
				const float2 p2 = mul2(rs[nb + 8 + j], bc2(h[8 + j]), nz);
				const float2 p3 = mul2(rs[nb + 12 + j], bc2(h[12 + j]), nz);
				L[j] = add2(add2(p0, p1), add2(p2, p3));
			}
			orow[(size_t)per * p] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
		}
	}
}

// ---------------------------------------------------------------------------------------------
// resampler_up_kernel — interpolating ratios (q <= p <= kRsMaxP, 16 taps), e.g. the 65/48 of the multi-ARFCN receive path.
// resampler16_kernel above reads a private 16-sample window per output from shared memory (one data-pipe wavefront per
// output: 76 % of the LSU peak at a third of the HBM roofline).  Consecutive outputs of an interpolator look at windows
// that move by 0 or 1 sample, so here a thread produces THREE consecutive outputs of one polyphase period from one
// 18-sample window held in registers (6 loads per output instead of 16):
//   rows      a CTA stages a tile of periods; period P's row holds input samples P*q - 15 .. P*q + q - 1 (the 15-sample
//             halo is stored again per row) at an odd pitch, so that lanes = consecutive periods read any row offset
//             without bank conflicts
//   items     (residue group g = outputs 3g .. 3g+2 of the period, block of 32 periods): lanes = periods.  The window
//             offsets and the three tap sets depend only on g, i.e. they are warp uniform: offsets are immediates after a
//             uniform branch, taps come from the kernel's parameter block (constant bank), nothing per lane
//   outputs   go to a [period][p] tile in shared memory (odd pitch: conflict free) and leave as one contiguous,
//             coalesced run of the stream
// Same sse_conv_real16 tree per output on the packed pipe as resampler16_kernel: bit-identical to Resampler::rotate.
// ---------------------------------------------------------------------------------------------
constexpr int kRsMaxP = 96;
constexpr int kRsMaxGroups = (kRsMaxP + 2) / 3;
constexpr int kRsTile = 64; // periods per tile
struct ResampUpParams {
	const float *in;
	float *out;
	int in_stride, out_len, out_stride, n_streams, p, q;
	float negzero;
	int goff[kRsMaxGroups];	      // per residue group: off0 | d1 << 8 | d2 << 16 (window offsets, see above)
	float gtaps[kRsMaxGroups][48]; // the three tap sets of the group (a missing residue repeats the first)
};
__host__ __device__ inline int rs_up_in_pitch(int q) { return (q + 15) | 1; }
__host__ __device__ inline int rs_up_out_pitch(int p) { return p | 1; }
__host__ __device__ inline size_t rs_up_smem(int p, int q) { return (size_t)kRsTile * (2 * rs_up_in_pitch(q) + rs_up_out_pitch(p)) * sizeof(float2); }
// host: fills goff / gtaps from the resampler's partition filters taps[p][16]
inline void rs_up_fill(ResampUpParams &P, const float *taps)
{
	const int p = P.p, q = P.q, ng = (p + 2) / 3;
	for (int g = 0; g < ng; g++) {
		const int rho0 = 3 * g, r1 = rho0 + 1 < p ? rho0 + 1 : rho0, r2 = rho0 + 2 < p ? rho0 + 2 : rho0;
		const int off0 = (q * rho0) / p;
		P.goff[g] = off0 | (((q * r1) / p - off0) << 8) | (((q * r2) / p - off0) << 16);
		const int rr[3] = { rho0, r1, r2 };
		for (int o = 0; o < 3; o++)
			for (int k = 0; k < 16; k++) P.gtaps[g][16 * o + k] = taps[(size_t)((q * rr[o]) % p) * 16 + k];
	}
}

template <int D>
__device__ __forceinline__ float2 rs_tree(const float2 (&x)[18], const float *h, float2 nz)
{
	float2 L[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const float2 p0 = mul2(x[D + j], bc2(h[j]), nz);
		const float2 p1 = mul2(x[D + 4 + j], bc2(h[4 + j]), nz);
		const float2 p2 = mul2(x[D + 8 + j], bc2(h[8 + j]), nz);
		const float2 p3 = mul2(x[D + 12 + j], bc2(h[12 + j]), nz);
		L[j] = add2(add2(p0, p1), add2(p2, p3));
	}
	return add2(add2(L[0], L[1]), add2(L[2], L[3]));
}

__global__ void __launch_bounds__(256, 2)
resampler_up_kernel(const __grid_constant__ ResampUpParams P)
{
	extern __shared__ __align__(16) float2 rsu[];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int p = P.p, q = P.q;
	const int ipitch = rs_up_in_pitch(q), opitch = rs_up_out_pitch(p);
	float2 *rout = rsu + (size_t)2 * kRsTile * ipitch; // [kRsTile][opitch] behind the two input buffers [kRsTile][ipitch]
	const float2 nz = make_float2(P.negzero, P.negzero);
	const int periods_total = P.out_len / p;
	const int tiles_per_stream = (periods_total + kRsTile - 1) / kRsTile;
	const int total_tiles = P.n_streams * tiles_per_stream;
	const int ngroups = (p + 2) / 3;
	const int rowlen = q + 15;
	const unsigned rin_s = (unsigned)__cvta_generic_to_shared(rsu);
	// stage: row r <- input samples (per0 + r) * q - 15 .. + q - 1; warps walk rows, a lane keeps its columns (coalesced
	// 8-byte asynchronous copies; per copy one shared and one global address increment)
	auto issue = [&](int tile_, int buf_) {
		const int s_ = tile_ / tiles_per_stream, per0_ = (tile_ - s_ * tiles_per_stream) * kRsTile;
		const int np_ = min(kRsTile, periods_total - per0_);
		const float2 *src = reinterpret_cast<const float2 *>(P.in) + (size_t)s_ * P.in_stride + ((long)per0_ * q - 15) + (long)warp * q + lane;
		unsigned dst = rin_s + 8u * (unsigned)(buf_ * kRsTile * ipitch + warp * ipitch + lane);
		for (int r = warp; r < np_; r += 8, src += 8 * q, dst += 64u * (unsigned)ipitch) {
			for (int c = 0; c + lane < rowlen; c += 32)
				asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * (unsigned)c), "l"(src + c) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	int buf = 0;
	if ((int)blockIdx.x < total_tiles) issue(blockIdx.x, 0);
	for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, buf ^= 1) {
		const int s = tile / tiles_per_stream, per0 = (tile - s * tiles_per_stream) * kRsTile;
		const int np = min(kRsTile, periods_total - per0);
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads(); // this tile's rows have landed; the other input buffer and the output tile are free
		if (tile + (int)gridDim.x < total_tiles) issue(tile + gridDim.x, buf ^ 1);
		const float2 *rin = rsu + (size_t)buf * kRsTile * ipitch;
		// ---- items: residue group g (warp uniform) x block of 32 periods (lanes) ----
		const int nblk = (np + 31) >> 5;
		for (int g = warp; g < ngroups; g += 8) {
			const int go = P.goff[g];
			const int off0 = go & 255, d1 = (go >> 8) & 255, d2 = (go >> 16) & 255;
			const int rho0 = 3 * g;
			float h[48];
#pragma unroll
			for (int k = 0; k < 48; k++) h[k] = P.gtaps[g][k];
			const bool full = off0 + 18 <= rowlen;
			for (int blk = 0; blk < nblk; blk++) {
				const int r = blk * 32 + lane;
				if (r < np) {
					const float2 *row = rin + (size_t)r * ipitch + off0;
					float2 x[18];
					if (full) {
#pragma unroll
						for (int k = 0; k < 18; k++) x[k] = row[k];
					} else {
#pragma unroll
						for (int k = 0; k < 18; k++) x[k] = (off0 + k < rowlen) ? row[k] : make_float2(0.0f, 0.0f);
					}
					float2 *orow = rout + (size_t)r * opitch + rho0;
					orow[0] = rs_tree<0>(x, h, nz);
					if (rho0 + 1 < p) orow[1] = d1 ? rs_tree<1>(x, h + 16, nz) : rs_tree<0>(x, h + 16, nz);
					if (rho0 + 2 < p) orow[2] = d2 == 2 ? rs_tree<2>(x, h + 32, nz) : (d2 == 1 ? rs_tree<1>(x, h + 32, nz) : rs_tree<0>(x, h + 32, nz));
				}
			}
		}
		__syncthreads();
		// ---- the tile's outputs are one contiguous run of the stream ----
		float2 *dst = reinterpret_cast<float2 *>(P.out) + (size_t)s * P.out_stride + (size_t)per0 * p;
		if (opitch == p && ((np * p) & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
			const float4 *r4 = reinterpret_cast<const float4 *>(rout);
			float4 *d4 = reinterpret_cast<float4 *>(dst);
			for (int idx = tid; idx < (np * p) >> 1; idx += 256) d4[idx] = r4[idx];
		} else if (opitch == p) {
			for (int idx = tid; idx < np * p; idx += 256) dst[idx] = rout[idx];
		} else {
			for (int r = warp; r < np; r += 8)
				for (int c = lane; c < p; c += 32) dst[r * p + c] = rout[(size_t)r * opitch + c];
		}
	}
}

// ---- channelizer ----
// in: [n_blocks*block_len][m] wideband samples (time-major), hist_in: [m][L] previous tail per branch,
// out: [m][n_blocks*block_len].  grid.x tiles time by T = 32.
__global__ void __launch_bounds__(256)
channelizer_kernel(const float *__restrict__ in, const float *__restrict__ hist_in, float *__restrict__ out, long out_stride,
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
				reinterpret_cast<float2 *>(out)[(size_t)c * out_stride + t] = make_float2(ar, ai);
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

// ---------------------------------------------------------------------------------------------
// M = 64, 16-tap fast paths (BASELINE config 5: 64-ARFCN wideband stream)
// ---------------------------------------------------------------------------------------------
// The generic kernels above evaluate the branch transform as an O(M^2) DFT (16,384 FMA per time index at
// M = 64), which makes them FP32 bound at a few percent of the HBM roofline.  Here the transform is the
// 8 x 8 Cooley-Tukey split of fft8.cuh (two in-register 8-point passes joined through shared memory,
// about 1,100 lane operations per time index) and the branch FIRs run as sliding windows in registers
// (thread = one branch x 8 consecutive time indices: 23 LDS.64 and 128 packed FMAs for 8 outputs).  The
// branch sums pass through the transform, whose parity bar is the mathematical DFT at 1e-4 (FFTW is
// unpinned in the reference), so FMA contraction is allowed in the FIR.
// Tile = 32 time indices x 64 branches per CTA pass, 256 threads:
//   FIR role   col = tid & 63 (lanes walk the wideband columns: coalesced / conflict free), g = tid >> 6
//   FFT role   lane = time index, warp w = residue of the branch index mod 8 (pass A) / output channel mod 8 (pass B)
// Every shared row has an odd float2 pitch, so lanes = time and lanes = branch are both conflict free.
constexpr int kFbT = 32;		   // time indices per tile
constexpr int kFbRows = kFbT + 15;	   // staged rows / columns incl. the 15-sample FIR halo
constexpr int kFbPitch = kFbT + 1;	   // float2 pitch of the [64][32] transform tiles
constexpr int kFbVPitch = kFbRows + 2;	   // float2 pitch of the synthesis FIR input rows (49)
constexpr size_t kCh64Smem = ((size_t)kFbRows * 64 + 2 * 64 * kFbPitch) * sizeof(float2);
constexpr size_t kSy64Smem = ((size_t)64 * kFbPitch + 64 * kFbVPitch) * sizeof(float2);

__device__ __forceinline__ float2 fb_fma2(float2 a, float h, float2 c)
{
	unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rc = *reinterpret_cast<unsigned long long *>(&c), rd;
	float2 hh = make_float2(h, h);
	unsigned long long rb = *reinterpret_cast<unsigned long long *>(&hh);
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
	return *reinterpret_cast<float2 *>(&rd);
}

// pass A of the 64-point transform for column `lane`: a[i] = sample of branch w + 8i  ->  z[(w*8 + k1)][lane]
__device__ __forceinline__ void fb_pass_a(float2 (&a)[8], const float2 (&twa)[8], float2 *z, int w, int lane)
{
	fft8(a);
#pragma unroll
	for (int k1 = 0; k1 < 8; k1++)
		z[(w * 8 + k1) * kFbPitch + lane] = k1 == 0 ? a[0] : c_mul(a[k1], twa[k1]);
}

// in: [total_t][64] wideband samples (time-major), hist_in: [64][16] previous tail per branch,
// out: [64][out_stride] (total_t samples written per row)
__global__ void __launch_bounds__(256, 3)
channelizer64_kernel(const float *__restrict__ in, const float *__restrict__ hist_in, float *__restrict__ out, long out_stride, long total_t,
		     const float *__restrict__ sub, const float2 *__restrict__ tw)
{
	extern __shared__ __align__(16) float2 fsm[];
	float2 *xs = fsm;			  // [47][64] wideband rows t0-15 .. t0+31
	float2 *y = fsm + kFbRows * 64;		  // [64][33] branch FIR outputs
	float2 *z = y + 64 * kFbPitch;		  // [64][33] pass-A outputs
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int col = tid & 63, g = tid >> 6, r = 63 - col; // deinterleave: in[i*m + n] -> branch m-1-n (Channelizer.cpp:44-45)
	float h[16];
#pragma unroll
	for (int k = 0; k < 16; k++) h[k] = __ldg(&sub[r * 16 + k]);
	float2 twa[8];
#pragma unroll
	for (int k1 = 0; k1 < 8; k1++) twa[k1] = __ldg(&tw[(w * k1) & 63]);
	const float2 *xin = reinterpret_cast<const float2 *>(in);
	const float2 *hin = reinterpret_cast<const float2 *>(hist_in);
	const long ntiles = (total_t + kFbT - 1) / kFbT;
	// Interior tiles are one contiguous 24,064-byte chunk: it is brought in with 16-byte cp.async copies issued as soon as
	// the FIR stage of the PREVIOUS tile has consumed xs, so the copy runs under that tile's two transform passes.
	auto interior = [&](long tile_) { const long t0_ = tile_ * kFbT; return t0_ >= 15 && t0_ + kFbT <= total_t; };
	auto prefetch = [&](long tile_) {
		const float4 *src = reinterpret_cast<const float4 *>(xin + (tile_ * kFbT - 15) * 64);
		const unsigned dst = (unsigned)__cvta_generic_to_shared(xs);
#pragma unroll
		for (int i = 0; i < 6; i++) {
			const int idx = tid + 256 * i;
			if (idx < kFbRows * 32)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (unsigned)idx), "l"(src + idx) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	bool fetched = false; // the current tile's rows are already on their way
	for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const long t0 = tile * kFbT;
		if (!fetched) {
			__syncthreads(); // previous tile is done with xs
			if (interior(tile)) {
				prefetch(tile);
			} else {
				for (int idx = tid; idx < kFbRows * 64; idx += 256) {
					const int row = idx >> 6, c = idx & 63;
					const long t = t0 - 15 + row;
					float2 v = make_float2(0.0f, 0.0f);
					if (t < 0) v = hin[(63 - c) * 16 + (int)(16 + t)];
					else if (t < total_t) v = __ldg(&xin[t * 64 + c]);
					xs[idx] = v;
				}
			}
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();
		// ---- branch FIRs: y[r][8g + o] = sum_k xs[8g + o + k][col] * h[k] ----
		{
			float2 acc[8];
#pragma unroll
			for (int o = 0; o < 8; o++) acc[o] = make_float2(0.0f, 0.0f);
#pragma unroll
			for (int j = 0; j < 23; j++) {
				const float2 x = xs[(8 * g + j) * 64 + col];
#pragma unroll
				for (int o = 0; o < 8; o++)
					if (j - o >= 0 && j - o < 16) acc[o] = fb_fma2(x, h[j - o], acc[o]);
			}
#pragma unroll
			for (int o = 0; o < 8; o++) y[r * kFbPitch + 8 * g + o] = acc[o];
		}
		__syncthreads();
		// xs is consumed: start the next tile's copy, it lands while the transform passes run
		{
			const long nxt = tile + gridDim.x;
			fetched = nxt < ntiles && interior(nxt);
			if (fetched) prefetch(nxt);
		}
		// ---- pass A: 8-point transforms over i of branch r = w + 8i, twiddled ----
		{
			float2 a[8];
#pragma unroll
			for (int i = 0; i < 8; i++) a[i] = y[(w + 8 * i) * kFbPitch + lane];
			fb_pass_a(a, twa, z, w, lane);
		}
		__syncthreads();
		// ---- pass B: 8-point transforms over j -> channels k1 + 8*k2, k1 = w; coalesced channel rows ----
		{
			float2 c[8];
#pragma unroll
			for (int j = 0; j < 8; j++) c[j] = z[(j * 8 + w) * kFbPitch + lane];
			fft8(c);
			const long t = t0 + lane;
			if (t < total_t) {
#pragma unroll
				for (int k2 = 0; k2 < 8; k2++)
					reinterpret_cast<float2 *>(out)[(size_t)(w + 8 * k2) * out_stride + t] = c[k2];
			}
		}
	}
}

// in: [64][total_t] per-channel samples, tail_in: [64][16] previous input tail per channel, out: [total_t][64].
// Each CTA owns a contiguous span of tiles, so the transformed columns the FIRs reach back to (15 per
// branch) are carried in shared memory from tile to tile; only a CTA's first tile recomputes them.
__global__ void __launch_bounds__(256, 3)
synthesis64_kernel(const float *__restrict__ in, const float *__restrict__ tail_in, float *__restrict__ out, long total_t,
		   int tiles_per_cta, const float *__restrict__ sub, const float2 *__restrict__ tw)
{
	extern __shared__ __align__(16) float2 fsm[];
	float2 *z = fsm;		      // [64][33] pass-A outputs
	float2 *v = fsm + 64 * kFbPitch;      // [64][49] transformed columns t0-15 .. t0+31 per branch
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int r = tid & 63, g = tid >> 6;
	float h[16];
#pragma unroll
	for (int k = 0; k < 16; k++) h[k] = __ldg(&sub[r * 16 + k]);
	float2 twa[8];
#pragma unroll
	for (int k1 = 0; k1 < 8; k1++) twa[k1] = __ldg(&tw[(w * k1) & 63]);
	const float2 *xin = reinterpret_cast<const float2 *>(in);
	const float2 *tin = reinterpret_cast<const float2 *>(tail_in);
	const long ntiles = (total_t + kFbT - 1) / kFbT;
	const long tile_lo = (long)blockIdx.x * tiles_per_cta;
	const long tile_hi = tile_lo + tiles_per_cta < ntiles ? tile_lo + tiles_per_cta : ntiles;
	for (long tile = tile_lo; tile < tile_hi; tile++) {
		const long t0 = tile * kFbT;
		// columns to transform this pass: the halo (first tile of the span) and then the tile itself
		for (int pass = (tile == tile_lo ? 0 : 1); pass < 2; pass++) {
			const long tb = pass ? t0 : t0 - 32; // column `lane` <-> time tb + lane; halo pass uses lanes 17..31
			const int vcol = pass ? 15 : -17;    // v column of lane 0
			const long t = tb + lane;
			float2 a[8];
#pragma unroll
			for (int i = 0; i < 8; i++) {
				const int c = w + 8 * i;
				float2 x = make_float2(0.0f, 0.0f);
				if (t >= 0) { if (t < total_t) x = __ldg(&xin[(size_t)c * total_t + t]); }
				else if (t >= -16) x = tin[c * 16 + (int)(16 + t)];
				a[i] = x;
			}
			__syncthreads(); // z free (previous pass B done), v columns of the previous tile consumed
			fb_pass_a(a, twa, z, w, lane);
			__syncthreads();
			float2 c[8];
#pragma unroll
			for (int j = 0; j < 8; j++) c[j] = z[(j * 8 + w) * kFbPitch + lane];
			fft8(c);
			if (vcol + lane >= 0) {
#pragma unroll
				for (int k2 = 0; k2 < 8; k2++) v[(w + 8 * k2) * kFbVPitch + vcol + lane] = c[k2];
			}
		}
		__syncthreads();
		// ---- branch FIRs over time, written interleaved out[t*64 + r] (Synthesis.cpp:38-49) ----
		{
			float2 acc[8];
#pragma unroll
			for (int o = 0; o < 8; o++) acc[o] = make_float2(0.0f, 0.0f);
#pragma unroll
			for (int j = 0; j < 23; j++) {
				const float2 x = v[r * kFbVPitch + 8 * g + j];
#pragma unroll
				for (int o = 0; o < 8; o++)
					if (j - o >= 0 && j - o < 16) acc[o] = fb_fma2(x, h[j - o], acc[o]);
			}
#pragma unroll
			for (int o = 0; o < 8; o++) {
				const long t = t0 + 8 * g + o;
				if (t < total_t) reinterpret_cast<float2 *>(out)[t * 64 + r] = acc[o];
			}
		}
		__syncthreads();
		// carry the last 15 transformed columns over as the next tile's halo
		if (tile + 1 < tile_hi) {
			float2 keep[4];
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const int idx = tid + 256 * i; // 64 rows x 15 columns = 960 items
				if (idx < 960) keep[i] = v[(idx / 15) * kFbVPitch + 32 + idx % 15];
			}
			__syncthreads();
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const int idx = tid + 256 * i;
				if (idx < 960) v[(idx / 15) * kFbVPitch + idx % 15] = keep[i];
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// M = 4, 8, 16 with 16-tap branches (the reference's own multi-ARFCN configuration is M = 4, radioInterfaceMulti.cpp:42)
// ---------------------------------------------------------------------------------------------
// With few branches a 32-index time tile leaves most of a CTA idle between two barriers.  Here a tile is 8 * 256 / M
// time indices (512 at M = 4): the FIR role is thread = (branch, 8 consecutive time indices) with a sliding register
// window as in the M = 64 kernels, the transform role is thread = time index with the M x M twiddle products unrolled
// from registers, so the wideband side moves in contiguous chunks and every channel row is written 256 bytes per
// warp instruction.  Rows of M samples get one pad row every 8 (position t + t / 8): both roles then read and write
// shared memory conflict free.  FMA contraction is allowed for the same reason as at M = 64 (the DFT's parity bar).
template <int M> struct FbSmall {
	static constexpr int G = 256 / M;	     // time groups of the FIR role
	static constexpr int T = 8 * G;		     // time indices per tile
	static constexpr int XROWS = T + 15 + (T + 15) / 8 + 1;
	static constexpr int YROWS = T + T / 8;
	static constexpr size_t kChSmem = (size_t)(XROWS + YROWS) * M * sizeof(float2);
	static constexpr size_t kSySmem = (size_t)XROWS * M * sizeof(float2);
};

template <int M>
__global__ void __launch_bounds__(256)
channelizer_small_kernel(const float *__restrict__ in, const float *__restrict__ hist_in, float *__restrict__ out, long out_stride, long total_t,
			 const float *__restrict__ sub, const float2 *__restrict__ tw)
{
	using C = FbSmall<M>;
	extern __shared__ __align__(16) float2 fsm[];
	float2 *xs = fsm;		       // [XROWS][M] wideband rows t0-15 .. t0+T-1 (padded positions)
	float2 *y = fsm + C::XROWS * M;	       // [YROWS][M] branch FIR outputs
	const int tid = threadIdx.x;
	const int col = tid % M, g = tid / M, r = M - 1 - col; // deinterleave: in[i*m + n] -> branch m-1-n (Channelizer.cpp:44-45)
	float h[16];
#pragma unroll
	for (int k = 0; k < 16; k++) h[k] = __ldg(&sub[r * 16 + k]);
	float2 w[M];
#pragma unroll
	for (int k = 0; k < M; k++) w[k] = __ldg(&tw[k]);
	const float2 *xin = reinterpret_cast<const float2 *>(in);
	const float2 *hin = reinterpret_cast<const float2 *>(hist_in);
	const long ntiles = (total_t + C::T - 1) / C::T;
	// Interior tiles (16-byte aligned input) are fetched with 16-byte cp.async copies, two samples of a row each, issued
	// as soon as the FIR stage of the previous tile has consumed xs: the copy runs under that tile's transform stage.
	const bool al16 = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
	auto interior = [&](long tile_) { const long t0_ = tile_ * C::T; return al16 && t0_ >= 15 && t0_ + C::T <= total_t; };
	auto prefetch = [&](long tile_) {
		const float4 *src = reinterpret_cast<const float4 *>(xin + (tile_ * C::T - 15) * M);
		const unsigned dst = (unsigned)__cvta_generic_to_shared(xs);
		for (int pc = tid; pc < (C::T + 15) * M / 2; pc += 256) {
			const int row = (2 * pc) / M, c = (2 * pc) % M;
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 8u * (unsigned)((row + (row >> 3)) * M + c)), "l"(src + pc)
				     : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	bool fetched = false;
	for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const long t0 = tile * C::T;
		if (!fetched) {
			__syncthreads();
			if (interior(tile)) {
				prefetch(tile);
			} else {
				for (int idx = tid; idx < (C::T + 15) * M; idx += 256) {
					const int row = idx / M, c = idx % M;
					const long t = t0 - 15 + row;
					float2 v = make_float2(0.0f, 0.0f);
					if (t < 0) v = hin[(M - 1 - c) * 16 + (int)(16 + t)];
					else if (t < total_t) v = __ldg(&xin[t * M + c]);
					xs[(row + (row >> 3)) * M + c] = v;
				}
			}
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();
		{
			float2 acc[8];
#pragma unroll
			for (int o = 0; o < 8; o++) acc[o] = make_float2(0.0f, 0.0f);
#pragma unroll
			for (int j = 0; j < 23; j++) {
				const float2 x = xs[(9 * g + j + (j >> 3)) * M + col];
#pragma unroll
				for (int o = 0; o < 8; o++)
					if (j - o >= 0 && j - o < 16) acc[o] = fb_fma2(x, h[j - o], acc[o]);
			}
#pragma unroll
			for (int o = 0; o < 8; o++) y[(9 * g + o) * M + r] = acc[o];
		}
		__syncthreads();
		{
			const long nxt = tile + gridDim.x; // xs is consumed: the next tile's copy lands during the transform stage
			fetched = nxt < ntiles && interior(nxt);
			if (fetched) prefetch(nxt);
		}
		for (int tt = tid; tt < C::T; tt += 256) {
			const long t = t0 + tt;
			float2 yv[M];
#pragma unroll
			for (int q = 0; q < M; q++) yv[q] = y[(tt + (tt >> 3)) * M + q];
			if (t < total_t) {
#pragma unroll
				for (int c = 0; c < M; c++) {
					float ar = 0.0f, ai = 0.0f;
#pragma unroll
					for (int q = 0; q < M; q++) {
						const float2 ww = w[(q * c) % M];
						ar = fmaf(yv[q].x, ww.x, ar); ar = fmaf(-yv[q].y, ww.y, ar);
						ai = fmaf(yv[q].x, ww.y, ai); ai = fmaf(yv[q].y, ww.x, ai);
					}
					reinterpret_cast<float2 *>(out)[(size_t)c * out_stride + t] = make_float2(ar, ai);
				}
			}
		}
	}
}

template <int M>
__global__ void __launch_bounds__(256)
synthesis_small_kernel(const float *__restrict__ in, const float *__restrict__ tail_in, float *__restrict__ out, long total_t,
		       const float *__restrict__ sub, const float2 *__restrict__ tw)
{
	using C = FbSmall<M>;
	extern __shared__ __align__(16) float2 fsm[];
	float2 *vs = fsm; // [XROWS][M] transformed columns t0-15 .. t0+T-1 (padded positions)
	const int tid = threadIdx.x;
	const int r = tid % M, g = tid / M;
	float h[16];
#pragma unroll
	for (int k = 0; k < 16; k++) h[k] = __ldg(&sub[r * 16 + k]);
	float2 w[M];
#pragma unroll
	for (int k = 0; k < M; k++) w[k] = __ldg(&tw[k]);
	const float2 *xin = reinterpret_cast<const float2 *>(in);
	const float2 *tin = reinterpret_cast<const float2 *>(tail_in);
	const long ntiles = (total_t + C::T - 1) / C::T;
	for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const long t0 = tile * C::T;
		__syncthreads();
		// forward DFT across channels for every column of the tile and its 15-column halo (recomputed: 3 % at M = 4)
		for (int j = tid; j < C::T + 15; j += 256) {
			const long t = t0 - 15 + j;
			float2 x[M];
#pragma unroll
			for (int c = 0; c < M; c++) {
				x[c] = make_float2(0.0f, 0.0f);
				if (t >= 0) { if (t < total_t) x[c] = __ldg(&xin[(size_t)c * total_t + t]); }
				else x[c] = tin[c * 16 + (int)(16 + t)];
			}
			float2 *vrow = vs + (j + (j >> 3)) * M;
#pragma unroll
			for (int q = 0; q < M; q++) {
				float ar = 0.0f, ai = 0.0f;
#pragma unroll
				for (int c = 0; c < M; c++) {
					const float2 ww = w[(q * c) % M];
					ar = fmaf(x[c].x, ww.x, ar); ar = fmaf(-x[c].y, ww.y, ar);
					ai = fmaf(x[c].x, ww.y, ai); ai = fmaf(x[c].y, ww.x, ai);
				}
				vrow[q] = make_float2(ar, ai);
			}
		}
		__syncthreads();
		// branch FIRs over time, written interleaved out[t*M + r] (Synthesis.cpp:38-49)
		{
			float2 acc[8];
#pragma unroll
			for (int o = 0; o < 8; o++) acc[o] = make_float2(0.0f, 0.0f);
#pragma unroll
			for (int j = 0; j < 23; j++) {
				const float2 x = vs[(9 * g + j + (j >> 3)) * M + r];
#pragma unroll
				for (int o = 0; o < 8; o++)
					if (j - o >= 0 && j - o < 16) acc[o] = fb_fma2(x, h[j - o], acc[o]);
			}
#pragma unroll
			for (int o = 0; o < 8; o++) {
				const long t = t0 + 8 * g + o;
				if (t < total_t) reinterpret_cast<float2 *>(out)[t * M + r] = acc[o];
			}
		}
	}
}

template <int M>
static void launch_channelizer_small(int sm_count, cudaStream_t st, const float *in, const float *hist, float *out, long out_stride,
				     long total_t, const float *sub, const float2 *tw)
{
	using C = FbSmall<M>;
	const long ntiles = (total_t + C::T - 1) / C::T;
	const int grid = (int)std::max<long>(1, std::min<long>(ntiles, (long)sm_count * 4));
	channelizer_small_kernel<M><<<grid, 256, C::kChSmem, st>>>(in, hist, out, out_stride, total_t, sub, tw);
}
template <int M>
static void launch_synthesis_small(int sm_count, cudaStream_t st, const float *in, const float *tail, float *out, long total_t,
				   const float *sub, const float2 *tw)
{
	using C = FbSmall<M>;
	const long ntiles = (total_t + C::T - 1) / C::T;
	const int grid = (int)std::max<long>(1, std::min<long>(ntiles, (long)sm_count * 4));
	synthesis_small_kernel<M><<<grid, 256, C::kSySmem, st>>>(in, tail, out, total_t, sub, tw);
}

__global__ void synthesis_tail_kernel(const float *__restrict__ in, float *__restrict__ tail_out, int m, int L, long total_t)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m * L) return;
	const int c = i / L, k = i % L;
	reinterpret_cast<float2 *>(tail_out)[(size_t)c * L + k] = reinterpret_cast<const float2 *>(in)[(size_t)c * total_t + (total_t - L + k)];
}

} // namespace trxb200
