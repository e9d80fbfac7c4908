// resampler_pq.cu — Resampler::rotate (Resampler.cpp:131-166) for the two ratios of the multi-ARFCN radio interface
// (radioInterfaceMulti.cpp:168-174: 65/48 on the receive side, 48/65 on the transmit side), 16 taps per path.
//
// resampler16_kernel (filterbank.cu) reads a private 16-sample window per output from shared memory: 32 data-pipe
// wavefronts per output, 76 % of the LSU peak at a third of the HBM roofline.  Consecutive outputs of one polyphase
// period look at windows that move by 0, 1 or 2 samples, so here a thread produces R CONSECUTIVE outputs of a period from
// one window of at most 21 samples held in registers (4 loads per output instead of 16), with the ratio a template
// parameter: window offsets are immediates; the taps are broadcast 16-byte loads from the CTA's copy in shared memory.
//   warp      = a tile of 32 consecutive periods of one stream, lanes = periods.  Nothing is shared between warps: no
//               block-wide barrier anywhere (resampler_up_kernel, the first attempt at register windows, spent its time in
//               two of them per tile).
//   input     the tile's span (32 q + 15 samples) is copied into the warp's rows [period][q] at an odd pitch (cp.async, 8
//               bytes per lane, coalesced): lane r reads row r + 1 at immediate offsets, conflict free; the 15 samples of
//               history in front of a period are the tail of the row above
//   items     residue group g = outputs R g .. R g + R - 1 of the period (the loop over g is unrolled: offsets, window
//               length and tap addresses are compile-time); sse_conv_real16's tree per output on the packed FP32 pipe,
//               products written fma(x, h, -0) with the -0 passed at run time (detect.cu) - bit-identical to the reference
//   output    goes to a [period][outputs of a third / half of a period] tile in shared memory (odd pitch) and leaves as contiguous runs
//               of the stream; the next tile's input is requested before the last half is written out
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

template <int P, int Q>
struct RsPqParams {
	const float *in; // per stream: 15 samples of history in front of in + s * in_stride (as resampler16_kernel)
	float *out;
	int in_stride, out_len, out_stride, n_streams;
	float negzero;
	float tp[P][16]; // taps of output residue rho = the partition filter of path (Q * rho) % P
};

template <int P, int Q, int R, int NCH>
struct RsPqGeom {
	static_assert(P % R == 0, "whole residue groups");
	static constexpr int groups = P / R;
	static constexpr int gchunk = (groups + NCH - 1) / NCH; // groups per output chunk
	static constexpr int in_pitch = Q | 1;
	static constexpr int out_pitch = (gchunk * R) | 1;
	static constexpr size_t warp_bytes = ((size_t)33 * in_pitch + (size_t)32 * out_pitch) * sizeof(float2);
	static constexpr size_t hdr_bytes = (size_t)P * 16 * sizeof(float);
	static constexpr int wmax = 16 + (Q * (R - 1) + P - 1) / P; // samples the windows of R consecutive outputs span, at most
	static_assert(wmax <= 24, "register window");
};

template <int P, int Q, int R, int NCH>
__global__ void __launch_bounds__(384, 1)
resampler_pq_kernel(const __grid_constant__ RsPqParams<P, Q> M)
{
	using G = RsPqGeom<P, Q, R, NCH>;
	extern __shared__ __align__(16) unsigned char rpq_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	// the CTA's copy of the taps in front of the warps' tiles: read as broadcast LDS.128 (straight from the parameter block they
	// arrive through the uniform datapath's constant loads, 317 of them per tile at the latency of the second-level constant
	// cache each - the 4 KB of taps sweep past the first level - and the kernel ran at 0.19 instructions per cycle and warp)
	const float4 *__restrict__ tp4 = reinterpret_cast<const float4 *>(rpq_raw); // [P][4]
	for (int k = threadIdx.x; k < P * 16; k += blockDim.x) reinterpret_cast<float *>(rpq_raw)[k] = M.tp[k >> 4][k & 15];
	__syncthreads();
	float2 *sin = reinterpret_cast<float2 *>(rpq_raw + G::hdr_bytes + (size_t)warp * G::warp_bytes); // [33][in_pitch]: row j = period per0 - 1 + j
	float2 *ost = sin + 33 * G::in_pitch;							 // [32][out_pitch]
	const unsigned sin_s = (unsigned)__cvta_generic_to_shared(sin);
	const float2 nz = make_float2(M.negzero, M.negzero);
	const int periods_total = M.out_len / P;
	const int tiles_per_stream = (periods_total + 31) >> 5;
	const long total_tiles = (long)M.n_streams * tiles_per_stream;
	const long tstep = (long)gridDim.x * wpb;

	auto issue = [&](long tile_) {
		const int s_ = (int)(tile_ / tiles_per_stream), per0_ = (int)(tile_ - (long)s_ * tiles_per_stream) << 5;
		const int np_ = min(32, periods_total - per0_);
		const float2 *src = reinterpret_cast<const float2 *>(M.in) + (size_t)s_ * M.in_stride + ((long)per0_ * Q - 15);
		const int cnt = np_ * Q + 15;
		// sample idx of the span sits u = idx + Q - 15 samples behind the start of row 0: row u / Q, column u % Q
		static_assert(Q > 32, "a step of 32 samples crosses at most one row boundary");
		int col = lane + Q - 15;
		unsigned dsts = sin_s + 8u * (unsigned)col;
		if (col >= Q) { col -= Q; dsts += 8u * (unsigned)(G::in_pitch - Q); }
		for (int idx = lane; idx < cnt; idx += 32) {
			asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dsts), "l"(src + idx) : "memory");
			col += 32; dsts += 256u;
			if (col >= Q) { col -= Q; dsts += 8u * (unsigned)(G::in_pitch - Q); }
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};

	long tile = (long)blockIdx.x * wpb + warp;
	if (tile < total_tiles) issue(tile);
	for (; tile < total_tiles; tile += tstep) {
		const int s = (int)(tile / tiles_per_stream), per0 = (int)(tile - (long)s * tiles_per_stream) << 5;
		const int np = min(32, periods_total - per0);
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
		const float2 *__restrict__ xrow = sin + (lane + 1) * G::in_pitch; // sample c of the lane's period: xrow[c], c >= 0; xrow[c - (in_pitch - Q)], c < 0
		float2 *dst = reinterpret_cast<float2 *>(M.out) + (size_t)s * M.out_stride + (size_t)per0 * P;
#pragma unroll
		for (int ch = 0; ch < NCH; ch++) {
			const int g_lo = ch * G::gchunk, g_hi = (ch + 1) * G::gchunk < G::groups ? (ch + 1) * G::gchunk : G::groups;
			float2 *__restrict__ orow = ost + lane * G::out_pitch;
			// the window of group g + 1 is requested before group g is evaluated, and a group's R outputs are stored together
			// behind its arithmetic (stores into the output tile would otherwise fence the loads of what follows)
			float2 xn[G::wmax];
			auto load_window = [&](int g) {
				const int w0 = (Q * (R * g)) / P;			 // window start of the group's first output (samples behind the period start)
				const int W = 16 + (Q * (R * g + R - 1)) / P - w0; // samples the R windows span
#pragma unroll
				for (int k = 0; k < G::wmax; k++) {
					const int c = w0 - 15 + k;
					if (k < W) xn[k] = xrow[c < 0 ? c - (G::in_pitch - Q) : c];
				}
			};
			load_window(g_lo);
#pragma unroll
			for (int g = g_lo; g < g_hi; g++) {
				const int w0 = (Q * (R * g)) / P;
				float2 x[G::wmax];
#pragma unroll
				for (int k = 0; k < G::wmax; k++) x[k] = xn[k];
				if (g + 1 < g_hi) load_window(g + 1);
				float2 y[R];
#pragma unroll
				for (int o = 0; o < R; o++) {
					const int rho = R * g + o, d = (Q * rho) / P - w0;
					float2 L[4];
					const float4 t0 = tp4[4 * rho], t1 = tp4[4 * rho + 1], t2 = tp4[4 * rho + 2], t3 = tp4[4 * rho + 3];
					const float h[16] = { t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w, t3.x, t3.y, t3.z, t3.w };
#pragma unroll
					for (int j = 0; j < 4; j++) {
						const float2 p0 = mul2(x[d + j], bc2(h[j]), nz);
						const float2 p1 = mul2(x[d + 4 + j], bc2(h[4 + j]), nz);
						const float2 p2 = mul2(x[d + 8 + j], bc2(h[8 + j]), nz);
						const float2 p3 = mul2(x[d + 12 + j], bc2(h[12 + j]), nz);
						L[j] = add2(add2(p0, p1), add2(p2, p3));
					}
					y[o] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
				}
#pragma unroll
				for (int o = 0; o < R; o++) orow[R * (g - g_lo) + o] = y[o];
			}
			__syncwarp();
			// the input rows are free once the last chunk has been evaluated: the next tile's span is on its way while this
			// chunk leaves
			if (ch == NCH - 1 && tile + tstep < total_tiles) issue(tile + tstep);
			// ---- the chunk's outputs: np runs of nco <= 32 samples, P apart in the stream; a lane keeps its column (the flat
			//      form - 32 consecutive elements per step, row and column by division - was 20 % of the kernel's instructions) ----
			const int c0 = R * g_lo, nco = R * (g_hi - g_lo);
			static_assert(G::gchunk * R <= 32, "one lane per output of a chunk");
			if (lane < nco) {
				const float2 *so = ost + lane;
				float2 *dg = dst + c0 + lane;
				int r = 0;
				for (; r + 4 <= np; r += 4) {
					const float2 v0 = so[r * G::out_pitch], v1 = so[(r + 1) * G::out_pitch], v2 = so[(r + 2) * G::out_pitch], v3 = so[(r + 3) * G::out_pitch];
					dg[(size_t)r * P] = v0; dg[(size_t)(r + 1) * P] = v1; dg[(size_t)(r + 2) * P] = v2; dg[(size_t)(r + 3) * P] = v3;
				}
				for (; r < np; r++) dg[(size_t)r * P] = so[r * G::out_pitch];
			}
			__syncwarp();
		}
	}
}

} // namespace trxb200
