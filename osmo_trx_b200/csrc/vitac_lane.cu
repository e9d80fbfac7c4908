// vitac_lane.cu — the grgsm_vitac MLSE (Transceiver52M/grgsm_vitac/) with LANE = BURST.
//
// vitac_kernel (vitac.cu) gives a warp to a pair of bursts and spreads each stage over the lanes: 59 correlation
// windows, 148 matched-filter outputs, 2 x 16 trellis states.  Two of its stages are serial per burst - the running
// window energy and the traceback run on ONE lane while 31 wait (a fifth of the kernel's instructions) - the others pay
// for shuffles, ballots, transposed staging and index arithmetic: 4,760 warp instructions per burst, issue bound.
// Here every thread equalises its own burst, exactly as the reference's scalar code does (which also makes the float
// order trivially the reference's): no lane is ever idle, nothing crosses lanes, and the 16 add-compare-select updates
// of a trellis step are 16 independent chains per thread.  About 1,500 warp instructions per burst.
//
//   CIR search   windows w = r + 4m of one residue r share all but one of their TLEN samples (taps are 4 samples
//                apart): a register window slides, ONE new sample is loaded per window (get_chan_imp_resp :183-235,
//                correlate_sequence :148-156); correlations and powers go to shared memory [window][lane]
//   energy       the running 20-window sum and its first maximum, per thread, in the reference's order
//   rhh, inc     autocorrelation of the 20 taps at lags 0, 4 .. 16 (:159-166, 93-95), increments (viterbi_detector.cc:93-100)
//   MF + ACS     per trellis step: four new samples into a 20-sample register window, the matched-filter component the
//                step consumes (mafi :168-181: imaginary part on even steps, real on odd), 16 state updates with
//                compile-time signs; the step's decisions (c2 > c1, c2 < c1 per state) are one word in shared memory
//                [step][lane]
//   traceback    per thread over its words (:371-391); decisions leave as 4-byte stores
// Rows are read straight from global memory, each thread walking its own row: a 32-byte sector serves four
// consecutive samples of a thread out of L1.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

constexpr int kVlWarps = 4; // warps per CTA
// shared memory per warp: max(correlations + powers [nwin][32] x 12 B, decision words [N][32] x 4 B)
__host__ __device__ inline size_t vl_warp_bytes(int nwin, int N)
{
	const size_t a = (size_t)nwin * 32 * 12, b = (size_t)N * 32 * 4;
	return ((a > b ? a : b) + 15) & ~(size_t)15;
}

// correlate_sequence for every search window, sliding register window per residue class
template <int TLEN, bool LANE_SEQ>
__device__ __forceinline__ void vl_search(const float2 *__restrict__ x0, const float2 *__restrict__ tseq, int nwin, float2 *cb, float *pw)
{
	float2 sq[LANE_SEQ ? TLEN : 1];
	if (LANE_SEQ) {
#pragma unroll
		for (int ii = 0; ii < TLEN; ii++) sq[ii] = tseq[ii];
	}
	const float ftlen = (float)TLEN;
	for (int r = 0; r < 4 && r < nwin; r++) {
		const float2 *x = x0 + r;
		float2 win[TLEN];
#pragma unroll
		for (int ii = 1; ii < TLEN; ii++) win[ii] = __ldg(&x[4 * (ii - 1)]);
		float2 nxt = __ldg(&x[4 * (TLEN - 1)]); // the one new sample of the next window, requested a whole window ahead
		for (int w = r, m = 0; w < nwin; w += 4, m++) {
#pragma unroll
			for (int ii = 0; ii < TLEN - 1; ii++) win[ii] = win[ii + 1];
			win[TLEN - 1] = nxt;
			if (w + 4 < nwin) nxt = __ldg(&x[4 * (m + TLEN)]);
			float rr = 0.0f, ri = 0.0f;
#pragma unroll
			for (int ii = 0; ii < TLEN; ii++) {
				const float2 s = LANE_SEQ ? sq[ii] : tseq[ii], v = win[ii];
				rr = fa(rr, fs(fm(s.x, v.x), fm(s.y, v.y)));
				ri = fa(ri, fa(fm(s.x, v.y), fm(s.y, v.x)));
			}
			const float2 c = make_float2(rr / ftlen, -ri / ftlen);
			cb[w * 32] = c;
			const float a = cabs_ref(c);
			pw[w * 32] = (float)((double)a * (double)a); // std::pow(abs(c), 2) evaluated in double
		}
	}
}

// one add-compare-select step for all 16 states; IMAG: the step consumes the imaginary matched-filter output
template <bool IMAG>
__device__ __forceinline__ unsigned vl_acs(const float (&o)[16], float (&nw)[16], float x, const float (&inc)[8])
{
	unsigned gt = 0u, lt = 0u;
#pragma unroll
	for (int s = 0; s < 16; s++) {
		const int pp = s >> 1, Ap = (pp ^ 2) & 7;
		const bool odd = s & 1;
		float c1, c2;
		if (IMAG) {
			// even s: c1 = old[p] + x - inc[A[p]], c2 = old[p+8] + x + inc[7-A[p]]; odd s: signs reversed
			const float sx = odd ? -x : x;
			c1 = fa(fa(o[pp], sx), odd ? inc[Ap] : -inc[Ap]);
			c2 = fa(fa(o[pp + 8], sx), odd ? -inc[7 - Ap] : inc[7 - Ap]);
		} else {
			// even s: c1 = old[p] - x - inc[7-p], c2 = old[p+8] - x + inc[p]; odd s: signs reversed
			const float sx = odd ? x : -x;
			c1 = fa(fa(o[pp], sx), odd ? inc[7 - pp] : -inc[7 - pp]);
			c2 = fa(fa(o[pp + 8], sx), odd ? -inc[pp] : inc[pp]);
		}
		nw[s] = (c2 < c1) ? c1 : c2; // the reference tests the sign of c2 - c1: a float difference has the sign of the comparison
		gt |= (c2 > c1 ? 1u : 0u) << s;
		lt |= (c2 < c1 ? 1u : 0u) << (16 + s);
	}
	return gt | lt;
}

// matched-filter component of one trellis step over the taps t < LIM of the window (LIM < 20 only at the burst's end)
template <bool IMAG, int LIM>
__device__ __forceinline__ float vl_mf(const float2 (&win)[20], const float2 (&cir)[20])
{
	float acc = 0.0f;
#pragma unroll
	for (int t = 0; t < LIM; t++) {
		const float2 v = win[t], c = cir[t];
		if (IMAG) acc = fa(acc, fa(fm(v.x, c.y), fm(v.y, c.x)));
		else acc = fa(acc, fs(fm(v.x, c.x), fm(v.y, c.y)));
	}
	return acc;
}

__device__ __forceinline__ void vl_prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

template <int TLEN, bool LANE_SEQ>
__global__ void __launch_bounds__(kVlWarps * 32, 2)
vitac_lane_kernel(VitacParams p)
{
	extern __shared__ __align__(16) unsigned char vl_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int N = p.is_ab == 1 ? 88 : 148;
	const int center = p.is_ab == 1 ? 13 : (p.is_ab == 2 ? 47 : 66);
	const int s0 = p.is_ab == 2 ? (center - 10) * kOSR : (center - 5) * kOSR + 1;
	const int s1 = p.is_ab == 2 ? (center + 30) * kOSR : (center + 5 + 5 + (p.is_ab ? p.max_delay : 0)) * kOSR;
	const int nwin = s1 - s0;
	unsigned char *wb = vl_raw + (size_t)warp * vl_warp_bytes(p.cir_in ? 0 : nwin, N);
	float2 *cb = reinterpret_cast<float2 *>(wb) + lane;			   // [nwin][32]
	float *pw = reinterpret_cast<float *>(wb + (size_t)nwin * 32 * 8) + lane; // [nwin][32]
	unsigned *words = reinterpret_cast<unsigned *>(wb) + lane;		   // [N][32], after the search is over

	const int ntiles = (p.n + 31) >> 5;
	for (int tile = blockIdx.x * kVlWarps + warp; tile < ntiles; tile += gridDim.x * kVlWarps) {
		const int b = tile * 32 + lane;
		const bool valid = b < p.n;
		const int bq = valid ? b : p.n - 1; // lanes beyond the batch repeat the last burst and store nothing
		const float2 *in = reinterpret_cast<const float2 *>(p.bufs) + (size_t)bq * p.stride + p.offset + (p.row_shift ? p.row_shift[bq] : 0);
		float2 cir[20];
		int st;
		__syncwarp();
		if (p.cir_in) {
			// detect_burst_nb / detect_burst_ab with the caller's channel estimate (:105-123): no search
#pragma unroll
			for (int k = 0; k < 20; k++) cir[k] = reinterpret_cast<const float2 *>(p.cir_in)[(size_t)bq * 20 + k];
			st = p.start_in ? max(p.clamp_lo, min(p.clamp_hi, p.start_in[bq])) : 0;
		} else {
			const float2 *tseq = p.is_ab == 2 ? &c_tab.vitac_sch[5]
						   : p.is_ab ? &c_tab.vitac_access[5] : &c_tab.vitac_norm[p.tsc ? (p.tsc[bq] > 8 ? 8 : p.tsc[bq]) : 0][5];
			vl_search<TLEN, LANE_SEQ>(in + s0, tseq, nwin, cb, pw);
			// ---- sliding 20-window energy, first maximum ----
			float ws = 0.0f;
			for (int i = 0; i < kCirLen; i++) ws = fa(ws, pw[i * 32]);
			float best = ws;
			int bi = 0;
			for (int i = kCirLen; i < nwin; i++) {
				ws = fa(ws, fs(pw[i * 32], pw[(i - kCirLen) * 32]));
				if (best < ws) { best = ws; bi = i - kCirLen + 1; }
			}
			float mc = 0.0f;
#pragma unroll
			for (int k = 0; k < 20; k++) {
				cir[k] = cb[(bi + k) * 32];
				const float a = cabs_ref(cir[k]);
				if (a > mc) mc = a;
			}
			st = max(p.clamp_lo, min(p.clamp_hi, s0 + bi - center * kOSR));
			if (valid) {
				p.corr_max[b] = mc;
				p.start[b] = st;
				if (p.cir) {
#pragma unroll
					for (int k = 0; k < 20; k++) reinterpret_cast<float2 *>(p.cir)[(size_t)b * 20 + k] = cir[k];
				}
			}
		}
		__syncwarp(); // the search area becomes the decision words
		// Every thread walks its own row, so a load instruction touches 32 lines and the register window only reaches one step
		// pair ahead: 39 % of the stall samples waited for those loads (ncu, long_scoreboard).  The rows are therefore asked for
		// in L2 well ahead of the walk: the first kilobyte of the burst now, a line every other step pair one kilobyte ahead in
		// the loop below, and the search region of the warp's NEXT tile while this one is equalised.
		{
			const float2 *xh = in + st;
#pragma unroll
			for (int k = 0; k < 9; k++) vl_prefetch_l2(xh + 16 * k);
			const int bn = (tile + (int)gridDim.x * kVlWarps) * 32 + lane;
			if (!p.cir_in && bn < p.n) {
				const float2 *inn = reinterpret_cast<const float2 *>(p.bufs) + (size_t)bn * p.stride + p.offset + (p.row_shift ? p.row_shift[bn] : 0) + s0;
				for (int k = 0; 16 * k < nwin + 4 * TLEN; k++) vl_prefetch_l2(inn + 16 * k);
			}
		}
		// ---- rhh[k] = conj(autocorr(cir)[4k]) (:159-166, 93-95); increments viterbi_detector.cc:93-100 ----
		float inc[8];
		{
			float rr[5], ri[5];
#pragma unroll
			for (int kk = 1; kk < 5; kk++) {
				const int k = 4 * kk;
				float ar = 0.0f, ai = 0.0f;
#pragma unroll
				for (int i = k; i < 20; i++) {
					const float2 a = cir[i], c = make_float2(cir[i - k].x, -cir[i - k].y);
					ar = fa(ar, fs(fm(a.x, c.x), fm(a.y, c.y)));
					ai = fa(ai, fa(fm(a.x, c.y), fm(a.y, c.x)));
				}
				rr[kk] = ar;
				ri[kk] = -ai;
			}
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const float a = (j & 1) ? ri[1] : -ri[1], bq2 = (j & 2) ? rr[2] : -rr[2], c = (j & 4) ? ri[3] : -ri[3];
				inc[j] = fa(fa(fa(a, bq2), c), rr[4]);
			}
		}
		// ---- matched filter + add-compare-select, two trellis steps per iteration ----
		float pa[16], pb[16];
#pragma unroll
		for (int s = 0; s < 16; s++) pa[s] = (s == p.start_state) ? 0.0f : (float)(-10e30);
		const float2 *x = in + st;
		float2 win[20];
#pragma unroll
		for (int t = 4; t < 20; t++) win[t] = __ldg(&x[t - 4]);
		// steps n <= N - 5 see all 20 taps (4 (N - n) >= 20); N is even, so whole pairs up to n = N - 6
		int n = 0;
		// the eight samples of a step pair are requested one pair (about 450 instructions) before they are used
		float2 nx[8];
#pragma unroll
		for (int t = 0; t < 8; t++) nx[t] = __ldg(&x[16 + t]);
#pragma unroll 1
		for (; n + 1 <= N - 5; n += 2) {
#pragma unroll
			for (int t = 0; t < 16; t++) win[t] = win[t + 4];
#pragma unroll
			for (int t = 0; t < 4; t++) win[16 + t] = nx[t];
			float2 hi4[4];
#pragma unroll
			for (int t = 0; t < 4; t++) hi4[t] = nx[4 + t];
			if (n + 3 <= N - 5) {
#pragma unroll
				for (int t = 0; t < 8; t++) nx[t] = __ldg(&x[4 * n + 24 + t]);
			}
			if ((n & 2) == 0) vl_prefetch_l2(&x[min(4 * n + 24 + 128, 4 * N + 12)]);
			words[n * 32] = vl_acs<true>(pa, pb, vl_mf<true, 20>(win, cir), inc);
#pragma unroll
			for (int t = 0; t < 16; t++) win[t] = win[t + 4];
#pragma unroll
			for (int t = 0; t < 4; t++) win[16 + t] = hi4[t];
			words[(n + 1) * 32] = vl_acs<false>(pb, pa, vl_mf<false, 20>(win, cir), inc);
		}
		// the last four steps (n = N - 4 .. N - 1): the filter runs off the end of the burst (mafi's break), 16, 12, 8, 4 taps
		{
#pragma unroll
			for (int t = 0; t < 16; t++) win[t] = win[t + 4];
			words[(N - 4) * 32] = vl_acs<true>(pa, pb, vl_mf<true, 16>(win, cir), inc);
#pragma unroll
			for (int t = 0; t < 12; t++) win[t] = win[t + 4];
			words[(N - 3) * 32] = vl_acs<false>(pb, pa, vl_mf<false, 12>(win, cir), inc);
#pragma unroll
			for (int t = 0; t < 8; t++) win[t] = win[t + 4];
			words[(N - 2) * 32] = vl_acs<true>(pa, pb, vl_mf<true, 8>(win, cir), inc);
#pragma unroll
			for (int t = 0; t < 4; t++) win[t] = win[t + 4];
			words[(N - 1) * 32] = vl_acs<false>(pb, pa, vl_mf<false, 4>(win, cir), inc);
		}
		// ---- best stop state of {4, 12} (viterbi_detector.cc:342-350), traceback (:371-391) ----
		unsigned state = (pa[12] > pa[4]) ? 12u : 4u;
		unsigned out_bit = 0u, real_imag = 0u; // N is even: the last step processed was a real one
		uint32_t *orow = reinterpret_cast<uint32_t *>(p.bits + (size_t)b * N);
		unsigned pack = 0u;
		for (int k = N - 1; k >= 0; k--) {
			const unsigned w = words[k * 32];
			const unsigned g = (w >> state) & 1u, l = (w >> (16 + state)) & 1u;
			const unsigned pos = (g != out_bit) ? l : g; // output[k] > 0
			pack = (pack << 8) | (pos ? 0x81u : 0x7fu);     // -127 : 127
			out_bit ^= real_imag ^ ((0x6666u >> state) & 1u);
			state = (state >> 1) + (g << 3);
			real_imag ^= 1u;
			if ((k & 3) == 0 && valid) orow[k >> 2] = pack;
		}
	}
}

} // namespace trxb200
