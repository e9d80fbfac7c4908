// detect.cu — batched burst detection (detectAnyBurst, sigProcLib.cpp:1926-1957) for sm_100a.
//
// Work decomposition (one warp owns a tile of 32 bursts, no block-level barriers):
//   phase A  (warp cooperates on one burst at a time, lanes = output samples)
//            4 sps -> 1 sps decimation of only the samples the correlator window needs
//            (downsampleBurst :1587-1601), correlation against the sync sequence (:1674) and the
//            candidate signal powers computeCI may need (:1622-1626).  Every output is an
//            independent dot product evaluated in the reference's float32 order, so spreading
//            outputs over lanes keeps results bit-identical.
//   phase C  (lanes = bursts) the inherently serial, comparison-driven part: argmax (:1120-1139),
//            edge gate (:1683), peak-to-average threshold (:1541-1571,1689), the 9-step early/late
//            TOA bisection over sinc-interpolated points (:1100-1118,1141-1186), C/I (:1608-1639),
//            amp = xcorr/gain, toa bookkeeping.  One lane per burst keeps all 32 lanes busy on
//            the long dependent chains; the correlation vectors live in shared memory laid out
//            [sample][lane] so that same-sample accesses are conflict free.
// The interpolation weights depend only on the tap distance and the position on the 1/512-symbol
// bisection grid, so the table-sinc (:990-998, double-precision index math) is folded on the host
// into interp_w[512][21] (tables.cpp); a weight is a single cached load.
// Multi-attempt types (EDGE -> TSC fall-through :1933-1941, EXT_RACH's three sequences :1793-1800)
// loop over attempts; only bursts that still need an attempt re-enter phase A.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

namespace {

constexpr float kClipThresh = 30000.0f; // CLIP_THRESH sigProcLib.cpp:49

// one decimated sample: sum_k x[4d-15+k] * g[k], sse_conv_real16 order (convolve_sse_3.c:188-264)
__device__ __forceinline__ float2 decimate_one(const float2 *__restrict__ x, int d)
{
	float pr[16], pi[16];
	const int base = 4 * d - 15;
#pragma unroll
	for (int k = 0; k < 16; k++) {
		const int idx = base + k;
		float2 v = make_float2(0.0f, 0.0f);
		if (idx >= 0)
			v = __ldg(&x[idx]);
		pr[k] = fm(v.x, c_tab.dnsamp[k]);
		pi[k] = fm(v.y, c_tab.dnsamp[k]);
	}
	float Lr[4], Li[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		Lr[j] = fa(fa(pr[j], pr[4 + j]), fa(pr[8 + j], pr[12 + j]));
		Li[j] = fa(fa(pi[j], pi[4 + j]), fa(pi[8 + j], pi[12 + j]));
	}
	return make_float2(fa(fa(Lr[0], Lr[1]), fa(Lr[2], Lr[3])), fa(fa(Li[0], Li[1]), fa(Li[2], Li[3])));
}

// one correlator output, sse_conv_cmplx_8n order (convolve_sse_3.c:462-537); hlen is 16, 40 or 64
__device__ __forceinline__ float2 correlate_one(const float2 *__restrict__ d, const float2 *__restrict__ h, int hlen)
{
	float Ar[4] = { 0, 0, 0, 0 }, Ai[4] = { 0, 0, 0, 0 }, Br[4] = { 0, 0, 0, 0 }, Bi[4] = { 0, 0, 0, 0 };
	for (int g = 0; g < hlen; g += 8) {
#pragma unroll
		for (int j = 0; j < 4; j++) {
			float2 xv = d[g + j], hv = h[g + j];
			Ar[j] = fa(Ar[j], fs(fm(hv.x, xv.x), fm(hv.y, xv.y)));
			Ai[j] = fa(Ai[j], fa(fm(hv.x, xv.y), fm(hv.y, xv.x)));
			xv = d[g + 4 + j];
			hv = h[g + 4 + j];
			Br[j] = fa(Br[j], fs(fm(hv.x, xv.x), fm(hv.y, xv.y)));
			Bi[j] = fa(Bi[j], fa(fm(hv.x, xv.y), fm(hv.y, xv.x)));
		}
	}
	float Lr[4], Li[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		Lr[j] = fa(Ar[j], Br[j]);
		Li[j] = fa(Ai[j], Bi[j]);
	}
	return make_float2(fa(fa(Lr[0], Lr[1]), fa(Lr[2], Lr[3])), fa(fa(Li[0], Li[1]), fa(Li[2], Li[3])));
}

struct Attempt {
	int seq;   // sequence id or -1
	int head;  // search symbols before target
	int start; // correlation start index (target - head - 1)
	int len;   // head + tail
	int rc_hit;
};

// window parameters per burst type and attempt (sigProcLib.cpp:1782-1924)
__device__ __forceinline__ Attempt make_attempt(int type, int tsc, int T, int attempt)
{
	Attempt a;
	a.seq = -1; a.head = 0; a.start = 0; a.len = 0; a.rc_hit = 0;
	switch (type) {
	case 5: // EDGE, then TSC
		if (attempt == 0) { a.seq = SEQ_EDGE + tsc; a.head = 6; a.start = 82 - 6 - 1; a.len = 6 + 6 + T; a.rc_hit = 5; }
		else if (attempt == 1) { a.seq = SEQ_MIDAMBLE + tsc; a.head = 10; a.start = 82 - 10 - 1; a.len = 10 + 6 + T; a.rc_hit = 1; }
		break;
	case 1:
		if (attempt == 0) { a.seq = SEQ_MIDAMBLE + tsc; a.head = 10; a.start = 71; a.len = 16 + T; a.rc_hit = 1; }
		break;
	case 2: // EXT_RACH
		if (attempt < 3) { a.seq = SEQ_RACH + attempt; a.head = 8; a.start = 48 - 8 - 1; a.len = 16 + T; a.rc_hit = 2; }
		break;
	case 3:
		if (attempt == 0) { a.seq = SEQ_RACH; a.head = 8; a.start = 39; a.len = 16 + T; a.rc_hit = 3; }
		break;
	case 6:
		if (attempt == 0) { a.seq = SEQ_DUMMY; a.head = 10; a.start = 71; a.len = 16 + T; a.rc_hit = 6; }
		break;
	default:
		break;
	}
	return a;
}

// interpolatePoint (sigProcLib.cpp:1100-1118) on a correlation vector stored [sample][lane]
__device__ __forceinline__ float2 interp_point(const float2 *corr, int lane, int len, float ix, const float *__restrict__ W)
{
	const int m = (int)floorf(ix);
	const int F = (int)((ix - (float)m) * 512.0f);
	int lo = m - 10, hi = m + 11;
	if (lo < 0) lo = 0;
	if ((unsigned)hi > (unsigned)(len - 1)) hi = len - 1;
	const float *w = W + F * 21;
	float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll 7
	for (int d = 0; d < 21; d++) {
		const int i = m - 10 + d;
		if (i >= lo && i < hi) {
			const float s = __ldg(&w[d]);
			const float2 v = corr[i * 32 + lane];
			acc.x = fa(acc.x, fm(v.x, s));
			acc.y = fa(acc.y, fm(v.y, s));
		}
	}
	return acc;
}

// the early and late points of one bisection step are exactly 2 samples apart, i.e. they sit on the same
// 1/512 grid position and use the same 21 weights: evaluate both in one pass (two independent chains)
__device__ __forceinline__ void interp_pair(const float2 *corr, int lane, int len, float early, const float *__restrict__ W,
					    float2 &pe, float2 &pl)
{
	const int m = (int)floorf(early);
	const int F = (int)((early - (float)m) * 512.0f);
	const int last = len - 1;
	int lo_e = max(m - 10, 0), hi_e = m + 11, lo_l = max(m - 8, 0), hi_l = m + 13;
	if ((unsigned)hi_e > (unsigned)last) hi_e = last;
	if ((unsigned)hi_l > (unsigned)last) hi_l = last;
	const float *w = W + F * 21;
	float er = 0.0f, ei = 0.0f, lr = 0.0f, li = 0.0f;
#pragma unroll 7
	for (int d = 0; d < 21; d++) {
		const int ie = m - 10 + d, il = ie + 2;
		const bool ve = ie >= lo_e && ie < hi_e, vl = il >= lo_l && il < hi_l;
		if (ve || vl) {
			const float s = __ldg(&w[d]);
			if (ve) {
				const float2 v = corr[ie * 32 + lane];
				er = fa(er, fm(v.x, s));
				ei = fa(ei, fm(v.y, s));
			}
			if (vl) {
				const float2 v = corr[il * 32 + lane];
				lr = fa(lr, fm(v.x, s));
				li = fa(li, fm(v.y, s));
			}
		}
	}
	pe = make_float2(er, ei);
	pl = make_float2(lr, li);
}

__device__ __forceinline__ bool near_tie(float a, float b)
{
	const float m = fmaxf(fabsf(a), fabsf(b));
	return fabsf(a - b) <= 4.0f * 1.1920929e-7f * m;
}

} // namespace

// Dynamic shared memory: per block the sync sequences + their SeqInfo (indexed per lane, so not read from
// __constant__), then per warp: corr [LMAX][32] float2, pwr [NDMAX][32] float (|dec|^2 of the decimated samples,
// for computeCI), dec [kGroup][NDMAX] float2, group parameters.
constexpr int kGroup = 4; // bursts decimated/correlated together in phase A (keeps the 32 lanes busy)

struct GroupSlot { int lb, seq_off, hlen, start, len, pad0, pad1, pad2; };

__host__ __device__ inline size_t detect_warp_bytes(int lmax, int ndmax)
{
	return (size_t)lmax * 32 * sizeof(float2) + (size_t)ndmax * 32 * sizeof(float) + (size_t)kGroup * ndmax * sizeof(float2) +
	       kGroup * sizeof(GroupSlot);
}
__host__ __device__ inline size_t detect_hdr_bytes()
{
	return SEQ_STORE * sizeof(float2) + ((SEQ_COUNT * sizeof(SeqInfo) + 15) & ~(size_t)15);
}

__global__ void __launch_bounds__(256, 3)
detect_kernel(DetectParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int warps_per_block = blockDim.x >> 5;
	float2 *sseq = reinterpret_cast<float2 *>(smem_raw);
	SeqInfo *sinfo = reinterpret_cast<SeqInfo *>(smem_raw + SEQ_STORE * sizeof(float2));
	unsigned char *base = smem_raw + detect_hdr_bytes() + detect_warp_bytes(p.lmax, p.ndmax) * warp;
	float2 *corr = reinterpret_cast<float2 *>(base);
	float *pwr = reinterpret_cast<float *>(base + (size_t)p.lmax * 32 * sizeof(float2));
	float2 *dec = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(pwr) + (size_t)p.ndmax * 32 * sizeof(float));
	GroupSlot *slot = reinterpret_cast<GroupSlot *>(dec + (size_t)kGroup * p.ndmax);
	const float *__restrict__ W = p.interp_w;

	for (int k = threadIdx.x; k < SEQ_STORE; k += blockDim.x) sseq[k] = c_tab.seq[k];
	for (int k = threadIdx.x; k < SEQ_COUNT; k += blockDim.x) sinfo[k] = c_tab.info[k];
	__syncthreads();

	const int ntiles = (p.n + 31) >> 5;
	for (int tile = blockIdx.x * warps_per_block + warp; tile < ntiles; tile += gridDim.x * warps_per_block) {
		const int b = tile * 32 + lane;
		const bool valid = b < p.n;
		int type = 0, tsc = 0, T = 0;
		if (valid) {
			type = p.type[b];
			tsc = p.tsc[b];
			T = p.max_toa[b];
		}
		int rc = 0;
		bool done = !valid;
		unsigned flags = 0;
		float2 amp = make_float2(0.0f, 0.0f);
		float toa = 0.0f, ci = 0.0f;
		int tsc_out = 0;
		bool clip = false;

		if (valid) {
			if ((type == 1 || type == 5) && tsc > 7) { rc = -3; done = true; tsc_out = 0; } // -SIGERR_UNSUPPORTED
			else if (type == 1 || type == 5) tsc_out = tsc;
			if (!done && (type == 1 || type == 2 || type == 3 || type == 5 || type == 6) && T > p.max_toa_bound) {
				rc = -1; done = true; // -SIGERR_BOUNDS: caller's bound was wrong
			}
			if (!(type == 1 || type == 2 || type == 3 || type == 5 || type == 6)) done = true; // "Invalid correlation type"
		}

		// optional clipping scan over the whole burst (maxAmplitude :1711-1722), warp per burst
		if (p.scan_clip) {
			unsigned need = __ballot_sync(0xffffffffu, valid && !done);
			for (unsigned m = need; m; m &= m - 1) {
				const int lb = __ffs(m) - 1;
				const float2 *x = reinterpret_cast<const float2 *>(p.bursts) + (size_t)(tile * 32 + lb) * p.stride;
				float mx = 0.0f;
				for (int i = lane; i < 625; i += 32) {
					const float2 v = __ldg(&x[i]);
					mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
				}
#pragma unroll
				for (int o = 16; o; o >>= 1)
					mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
				if (lane == lb)
					clip = mx > kClipThresh;
			}
			if (clip) flags |= 4u;
		}

		for (int attempt = 0; attempt < 3; attempt++) {
			Attempt at = make_attempt(type, tsc, T, attempt);
			if (!done && at.seq >= 0 && sinfo[at.seq].len + at.len - 1 > p.ndmax) {
				rc = -1; done = true; // window larger than trxb200_detect_config() promised: -SIGERR_BOUNDS
			}
			const bool need = !done && at.seq >= 0;
			const unsigned mask = __ballot_sync(0xffffffffu, need);
			if (!mask)
				break;
			const float2 *tile_x = reinterpret_cast<const float2 *>(p.bursts) + (size_t)tile * 32 * p.stride;

			// ---- phase A: kGroup bursts at a time, work items flattened over (burst, output) ----
			unsigned m = mask;
			while (m) {
				// lanes 0..kGroup-1 publish the parameters of the next bursts of the mask
				int ng = 0, ndpad = 0, lenpad = 0;
				__syncwarp();
#pragma unroll
				for (int g = 0; g < kGroup; g++) {
					if (m) {
						const int lb = __ffs(m) - 1;
						m &= m - 1;
						const int seq = __shfl_sync(0xffffffffu, at.seq, lb);
						const int start = __shfl_sync(0xffffffffu, at.start, lb);
						const int len = __shfl_sync(0xffffffffu, at.len, lb);
						const int hl = sinfo[seq].len;
						if (lane == 0) {
							GroupSlot gs;
							gs.lb = lb; gs.seq_off = sinfo[seq].off; gs.hlen = hl; gs.start = start; gs.len = len;
							gs.pad0 = gs.pad1 = gs.pad2 = 0;
							slot[g] = gs;
						}
						ndpad = max(ndpad, hl + len - 1);
						lenpad = max(lenpad, len);
						ng = g + 1;
					}
				}
				__syncwarp();
				// decimation of the samples the correlators need (+ their powers for computeCI)
				for (int it = lane; it < ng * ndpad; it += 32) {
					const int g = (it >= ndpad) + (it >= 2 * ndpad) + (it >= 3 * ndpad);
					const int j = it - g * ndpad;
					const GroupSlot gs = slot[g];
					if (j < gs.hlen + gs.len - 1) {
						const int d = gs.start - (gs.hlen - 1) + j;
						float2 v = make_float2(0.0f, 0.0f);
						if (d >= 0 && d < 156)
							v = decimate_one(tile_x + (size_t)gs.lb * p.stride, d);
						dec[g * p.ndmax + j] = v;
						pwr[j * 32 + gs.lb] = norm2(v);
					}
				}
				__syncwarp();
				// correlation
				for (int it = lane; it < ng * lenpad; it += 32) {
					const int g = (it >= lenpad) + (it >= 2 * lenpad) + (it >= 3 * lenpad);
					const int i = it - g * lenpad;
					const GroupSlot gs = slot[g];
					if (i < gs.len)
						corr[i * 32 + gs.lb] = correlate_one(dec + g * p.ndmax + i, sseq + gs.seq_off, gs.hlen);
				}
			}
			__syncwarp();

			// ---- phase C ----
			if (need) {
				const int len = at.len;
				const SeqInfo si = sinfo[at.seq];
				// fastPeakDetect
				float mx = 0.0f;
				int idx = -1;
				float2 pk = make_float2(0.0f, 0.0f);
				for (int i = 0; i < len; i++) {
					const float2 v = corr[i * 32 + lane];
					const float pwv = norm2(v);
					if (pwv > mx) { mx = pwv; idx = i; pk = v; }
				}
				float t = (float)idx;
				bool hit = !((t < 3.0f) || (t > (float)(len - 3)));
				if (hit) {
					// computePeakRatio (sps = 1)
					int num = 0;
					float avg = 0.0f;
					for (int i = 2; i <= 5; i++) {
						if (idx - i >= 0) { avg = fa(avg, norm2(corr[(idx - i) * 32 + lane])); num++; }
						if (idx + i < len) { avg = fa(avg, norm2(corr[(idx + i) * 32 + lane])); num++; }
					}
					float ratio = 0.0f;
					if (num >= 5) {
						const float rms = (float)((double)sqrtf(avg / (float)num) + 0.00001);
						ratio = sqrtf(norm2(pk)) / rms;
					}
					if (fabsf(ratio - p.thresh) < 1e-5f) flags |= 1u;
					if (ratio < p.thresh) hit = false;
				}
				if (hit) {
					// peakDetect: early/late bisection
					float early = t - 1.0f, late = t + 1.0f, incr = 0.5f;
#pragma unroll 1
					for (int it = 0; it < 9; it++) {
						float2 e, l;
						interp_pair(corr, lane, len, early, W, e, l); // late == early + 2 throughout (:1172)
						const float ne = norm2(e), nl = norm2(l);
						if (near_tie(ne, nl)) flags |= 2u;
						if (ne < nl) early += incr;
						else if (ne > nl) early -= incr;
						else break;
						incr *= 0.5f;
						late = early + 2.0f;
					}
					t = early + 1.0f;
					const float2 xc = interp_point(corr, lane, len, t, W);
					// computeCI
					const int N = si.len;
					const int rt = (int)roundf(t);
					const int ps = at.start + 1 - N + rt;
					if (ps < 0 || ps + N > 156 || rt < 0 || rt >= len) {
						ci = 0.0f;
					} else {
						// S = mean |burst[ps..ps+N)|^2, sequential (:1622-1626); pwr index j = dec index - d0 = rt + k
						float S = 0.0f;
						for (int k = 0; k < N; k++)
							S = fa(S, pwr[(rt + k) * 32 + lane]);
						S = S / (float)N;
						const float C = norm2(xc) / si.ci_den;
						ci = fm(3.0103f, log2f(C / fs(S, C)));
					}
					amp = cmul_exact(xc, make_float2(si.inv_gr, si.inv_gi));
					toa = fs(fs(t, si.toa), (float)at.head);
					rc = at.rc_hit;
					if (type == 2 || type == 3) tsc_out = attempt;
					done = true;
				}
			}
		}

		if (valid) {
			if (rc == 0 && clip) rc = -2; // -SIGERR_CLIP only when nothing was detected (:1764)
			p.rc[b] = rc;
			reinterpret_cast<float2 *>(p.amp)[b] = amp;
			p.toa[b] = toa;
			p.ci[b] = ci;
			p.tsc_out[b] = (uint8_t)tsc_out;
			if (p.flags) p.flags[b] = (uint8_t)flags;
		}
	}
}

} // namespace trxb200
