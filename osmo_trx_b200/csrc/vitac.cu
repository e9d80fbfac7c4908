// vitac.cu — batched grgsm_vitac MLSE equaliser (Transceiver52M/grgsm_vitac/) for sm_100a.
//
// Per burst: training-sequence CIR search (get_chan_imp_resp, grgsm_vitac.cpp:183-235), CIR
// autocorrelation -> rhh (:159-166, 93-95), matched filter (mafi :168-181), 16-state Viterbi with
// real/imaginary alternation and traceback (viterbi_detector.cc:63-392).
//
// Mapping: one warp works on a pair of bursts.  The data-parallel parts (59+ correlation windows,
// 148 matched-filter outputs) spread outputs over the 32 lanes, one burst after the other; each
// output keeps the reference's sequential float32 accumulation order, so path metrics — and with
// them every decision — are bit-identical.  Only the component the trellis consumes (imag on even
// symbols, real on odd ones, viterbi_detector.cc:118-121,228-230) is computed.  The add-compare-
// select runs with lane = (burst, state): predecessors old[s>>1] and old[(s>>1)+8] arrive by
// warp shuffle, and because detect_burst_generic only keeps the SIGN of the traceback output
// (grgsm_vitac.cpp:101-102) the 148x16 float trans_table is replaced by two ballot words per step
// (diff > 0, diff < 0) staged in shared memory for the serial traceback.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

struct VitacParams {
	const float *bufs;
	int stride, offset, n, is_ab;
	const uint8_t *tsc;
	int max_delay, clamp_lo, clamp_hi;
	int8_t *bits;
	int32_t *start;
	float *corr_max, *cir;
	int nwin_max;
	const float *cir_in;	  // complex[n][20] caller's channel estimates, or null: estimate here (get_*_imp_resp)
	const int32_t *start_in; // burst start per burst when cir_in is given
	const int32_t *row_shift = nullptr; // cir_in mode: per-burst sample offset added to the row base (the burst found by a long search); start is 0 then
	int start_state = 3;	  // viterbi_detector's start state (detect_burst_* default 3, grgsm_vitac.cpp:114-121)
	int lo, range, pitch; // staged part of each row: samples [lo, lo + range) relative to the burst; plane pitch of the window
};

namespace {

constexpr int kOSR = 4;
constexpr int kCirLen = 20; // CHAN_IMP_RESP_LENGTH * OSR

// std::abs(std::complex<float>) = hypotf, evaluated by glibc as (float)sqrt((double)x*x + (double)y*y)
__device__ __forceinline__ float cabs_ref(float2 c)
{
	return (float)sqrt((double)c.x * (double)c.x + (double)c.y * (double)c.y);
}

} // namespace

// plane pitch (in samples) of the staged window: >= range / 4 + 1 and = 4 (mod 16), so that a warp's 32 consecutive
// samples (8 per plane) and a warp's 32 samples at stride 4 (one plane) both fall on distinct banks
__host__ __device__ inline int vitac_pitch(int range) { int pp = (range + 3) / 4 + 1; while ((pp & 15) != 4) pp++; return pp; }
__host__ __device__ inline int vitac_warp_floats(int nwin_max, int pitch)
{
	return (8 * pitch + 2 * nwin_max + ((nwin_max + 1) & ~1) + 2 * 160 + 16 + 80 + 2 * 160 + 8 + 3) & ~3;
}

// shared memory per warp (floats): xs[4][pitch] complex | cb[2*nwin] | pw[nwin] | filt[2][160] | inc[2][8] | csel[2][20] complex |
// words[160][2] | misc[8]
//
// Each row is staged ONCE (coalesced 8-byte loads) into a transposed window, sample r in plane r & 3 at index r >> 2:
// the correlation windows (lanes = consecutive samples, taps 4 samples apart) and the matched filter (lanes = outputs
// 4 samples apart, taps consecutive) then read it conflict free, instead of fetching 32 different 32-byte sectors per
// matched-filter load from L1.
__global__ void __launch_bounds__(128)
vitac_kernel(VitacParams p)
{
	extern __shared__ __align__(16) float vsm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	const int P = p.pitch;
	const int per_warp = vitac_warp_floats(p.nwin_max, P);
	float *base = vsm + (size_t)warp * per_warp;
	float2 *xs = reinterpret_cast<float2 *>(base);
	float2 *cb = xs + 4 * P;
	float *pw = reinterpret_cast<float *>(cb + p.nwin_max);
	float *filt = pw + ((p.nwin_max + 1) & ~1); // [2][160], 8-byte aligned
	float *inc = filt + 320;	       // [2][8]
	float2 *csel = reinterpret_cast<float2 *>(inc + 16); // [2][20]: CIR taps arranged for the imaginary / real matched-filter outputs
	unsigned *words = reinterpret_cast<unsigned *>(base + (per_warp - 8 - 320)); // [160][2]: gt, lt (16-byte aligned)
	int *misc = reinterpret_cast<int *>(words + 320);

	// is_ab: 0 normal burst (get_norm_chan_imp_resp), 1 access burst (get_access_imp_resp), 2 SCH burst
	// (get_sch_chan_imp_resp :283-296: centre SYNC_POS + 5, ten symbols back, SYNC_SEARCH_RANGE forward, 54 symbols)
	const int N = p.is_ab == 1 ? 88 : 148;
	const int center = p.is_ab == 1 ? 13 : (p.is_ab == 2 ? 47 : 66);
	const int s0 = p.is_ab == 2 ? (center - 10) * kOSR : (center - 5) * kOSR + 1;
	const int s1 = p.is_ab == 2 ? (center + 30) * kOSR : (center + 5 + 5 + (p.is_ab ? p.max_delay : 0)) * kOSR;
	const int nwin = s1 - s0;
	const int tlen = p.is_ab == 1 ? 31 : (p.is_ab == 2 ? 54 : 16);
	const float ftlen = (float)tlen;

	const int npairs = (p.n + 1) >> 1;
	for (int pair = blockIdx.x * wpb + warp; pair < npairs; pair += gridDim.x * wpb) {
		__syncwarp();
		// the rows of the warp's NEXT pair are requested into L2 now: the staging loads below otherwise wait on HBM with
		// nothing else for this warp to do (21 % of the kernel's stall samples sat on them, profiles/r1x_vitac_summary.txt)
		{
			const int np_ = pair + gridDim.x * wpb;
			if (np_ < npairs) {
				for (int h = 0; h < 2; h++) {
					const int b = 2 * np_ + h;
					if (b < p.n) {
						const char *row = reinterpret_cast<const char *>(reinterpret_cast<const float2 *>(p.bufs) + (size_t)b * p.stride + p.offset + p.lo +
											       (p.row_shift ? p.row_shift[b] : 0));
						for (int o = lane * 128; o < p.range * 8; o += 32 * 128)
							asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
					}
				}
			}
		}
		for (int h = 0; h < 2; h++) {
			const int b = 2 * pair + h;
			if (b >= p.n) {
				for (int i = lane; i < 160; i += 32) filt[h * 160 + i] = 0.0f;
				if (lane < 8) inc[h * 8 + lane] = 0.0f;
				__syncwarp(); // the ACS stage reads these
				continue;
			}
			const float2 *in = reinterpret_cast<const float2 *>(p.bufs) + (size_t)b * p.stride + p.offset + p.lo + (p.row_shift ? p.row_shift[b] : 0);
			const float2 *tseq = p.is_ab == 2 ? &c_tab.vitac_sch[5]
						   : p.is_ab ? &c_tab.vitac_access[5] : &c_tab.vitac_norm[p.tsc ? (p.tsc[b] > 8 ? 8 : p.tsc[b]) : 0][5];
			// ---- stage the row ----
			for (int r0 = 0; r0 < p.range; r0 += 256) {
				float2 v[8];
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const int r = r0 + lane + 32 * k;
					v[k] = make_float2(0.0f, 0.0f);
					if (r < p.range) v[k] = __ldg(&in[r]);
				}
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const int r = r0 + lane + 32 * k;
					if (r < p.range) xs[(r & 3) * P + (r >> 2)] = v[k];
				}
			}
			__syncwarp();
			int bi = 0, st = 0;
			if (p.cir_in) {
				// detect_burst_nb / detect_burst_ab with the caller's channel estimate (:105-123): no search
				if (lane < kCirLen) cb[lane] = reinterpret_cast<const float2 *>(p.cir_in)[(size_t)b * kCirLen + lane];
				st = p.start_in ? max(p.clamp_lo, min(p.clamp_hi, p.start_in[b])) : 0;
				__syncwarp();
			} else {
			// ---- correlation per search window (correlate_sequence :148-156) ----
			for (int w = lane; w < nwin; w += 32) {
				const int r = s0 + w - p.lo;
				const float2 *x = xs + (r & 3) * P + (r >> 2); // taps 4 samples apart: consecutive in the plane
				float rr = 0.0f, ri = 0.0f;
				for (int ii = 0; ii < tlen; ii++) {
					const float2 s = tseq[ii], v = x[ii];
					rr = fa(rr, fs(fm(s.x, v.x), fm(s.y, v.y)));
					ri = fa(ri, fa(fm(s.x, v.y), fm(s.y, v.x)));
				}
				const float2 c = make_float2(rr / ftlen, -ri / ftlen);
				cb[w] = c;
				const float a = cabs_ref(c);
				pw[w] = (float)((double)a * (double)a); // std::pow(abs(c), 2) evaluated in double
			}
			__syncwarp();
			// ---- sliding 20-window energy, first maximum (serial, lane 0) ----
			if (lane == 0) {
				float ws = 0.0f;
				for (int i = 0; i < kCirLen; i++) ws = fa(ws, pw[i]);
				float best = ws;
				int bi = 0, idx = 0;
				for (int i = kCirLen; i < nwin; i++) {
					ws = fa(ws, fs(pw[i], pw[i - kCirLen]));
					idx++;
					if (best < ws) { best = ws; bi = idx; }
				}
				misc[h] = bi;
			}
			__syncwarp();
			bi = misc[h];
			st = s0 + bi - center * kOSR;
			st = max(p.clamp_lo, min(p.clamp_hi, st));
			// corr_max + CIR export
			{
				float a = (lane < kCirLen) ? cabs_ref(cb[bi + lane]) : 0.0f;
#pragma unroll
				for (int o = 16; o; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
				if (lane == 0) { p.corr_max[b] = a; p.start[b] = st; }
				if (p.cir && lane < kCirLen)
					reinterpret_cast<float2 *>(p.cir)[(size_t)b * kCirLen + lane] = cb[bi + lane];
			}
			}
			const float2 *cir = cb + bi;
			// ---- rhh[k] = conj(autocorr(cir)[4k]) (:159-166,93-95); increments viterbi_detector.cc:93-100 ----
			if (lane < 5) {
				const int k = 4 * lane;
				float ar = 0.0f, ai = 0.0f;
				for (int i = k; i < kCirLen; i++) {
					const float2 a = cir[i], c = make_float2(cir[i - k].x, -cir[i - k].y);
					ar = fa(ar, fs(fm(a.x, c.x), fm(a.y, c.y)));
					ai = fa(ai, fa(fm(a.x, c.y), fm(a.y, c.x)));
				}
				filt[h * 160 + 150 + lane] = ar;   // scratch: rhh real
				filt[h * 160 + 155 + lane] = -ai;  // rhh imag (conjugated)
			}
			__syncwarp();
			if (lane < 8) {
				const float r1i = filt[h * 160 + 156], r2r = filt[h * 160 + 152], r3i = filt[h * 160 + 158], r4r = filt[h * 160 + 154];
				const float a = (lane & 1) ? r1i : -r1i;
				const float bq = (lane & 2) ? r2r : -r2r;
				const float c = (lane & 4) ? r3i : -r3i;
				inc[h * 8 + lane] = fa(fa(fa(a, bq), c), r4r);
			}
			// ---- matched filter (mafi :168-181), only the component the trellis reads: even outputs the imaginary
			//      part v.x*c.y + v.y*c.x, odd outputs the real part v.x*c.x - v.y*c.y.  Output parity = lane parity, so
			//      each lane reads its taps from the arrangement (c.y, c.x) or (c.x, -c.y): one expression, no divergence
			//      (a - b and a + (-b) are the same float). ----
			if (lane < kCirLen) {
				const float2 c = cir[lane];
				csel[lane] = make_float2(c.y, c.x);
				csel[kCirLen + lane] = make_float2(c.x, -c.y);
			}
			__syncwarp();
			{
				const int rb = st - p.lo; // window position of the burst's first sample (warp uniform)
				const int e0 = rb & 3;
				const float2 *cs = csel + (lane & 1) * kCirLen;
				for (int nn = lane; nn < N; nn += 32) {
					const float2 *xq = xs + (rb >> 2) + nn; // sample rb + 4 nn + ii: plane (e0 + ii) & 3, index + (e0 + ii) >> 2
					const int lim = min(kCirLen, kOSR * (N - nn));
					float acc = 0.0f;
					if (lim == kCirLen) {
#pragma unroll
						for (int ii = 0; ii < kCirLen; ii++) {
							const int e = e0 + ii;
							const float2 v = xq[(e & 3) * P + (e >> 2)], c = cs[ii];
							acc = fa(acc, fa(fm(v.x, c.x), fm(v.y, c.y)));
						}
					} else {
						for (int ii = 0; ii < lim; ii++) {
							const int e = e0 + ii;
							const float2 v = xq[(e & 3) * P + (e >> 2)], c = cs[ii];
							acc = fa(acc, fa(fm(v.x, c.x), fm(v.y, c.y)));
						}
					}
					filt[h * 160 + nn] = acc;
				}
			}
			__syncwarp();
		}

		// ---- add-compare-select: lane = (burst h, state s) ----
		{
			const int h = lane >> 4, s = lane & 15, pp = s >> 1;
			const bool odd = s & 1;
			const int Ap = (pp ^ 2) & 7; // {2,3,0,1,6,7,4,5}
			const float *ic = inc + h * 8;
			const float i1I = odd ? ic[Ap] : -ic[Ap], i2I = odd ? -ic[7 - Ap] : ic[7 - Ap];
			const float i1R = odd ? ic[7 - pp] : -ic[7 - pp], i2R = odd ? -ic[pp] : ic[pp];
			float pm = (s == p.start_state) ? 0.0f : (float)(-10e30);
			const int src1 = (h << 4) + pp, src2 = src1 + 8;
			const float2 *f2 = reinterpret_cast<const float2 *>(filt + h * 160);
			// two trellis steps per iteration (N is even): the imaginary step, then the real one
			for (int k = 0; k < N; k += 2) {
				const float2 x = f2[k >> 1];
				uint4 wd;
				{
					const float o1 = __shfl_sync(0xffffffffu, pm, src1), o2 = __shfl_sync(0xffffffffu, pm, src2);
					const float sx = odd ? -x.x : x.x;
					const float c1 = fa(fa(o1, sx), i1I), c2 = fa(fa(o2, sx), i2I);
					// the reference tests the sign of c2 - c1; a float difference has the sign of the comparison
					pm = (c2 < c1) ? c1 : c2;
					wd.x = __ballot_sync(0xffffffffu, c2 > c1);
					wd.y = __ballot_sync(0xffffffffu, c2 < c1);
				}
				{
					const float o1 = __shfl_sync(0xffffffffu, pm, src1), o2 = __shfl_sync(0xffffffffu, pm, src2);
					const float sx = odd ? x.y : -x.y;
					const float c1 = fa(fa(o1, sx), i1R), c2 = fa(fa(o2, sx), i2R);
					pm = (c2 < c1) ? c1 : c2;
					wd.z = __ballot_sync(0xffffffffu, c2 > c1);
					wd.w = __ballot_sync(0xffffffffu, c2 < c1);
				}
				if (lane == 0) reinterpret_cast<uint4 *>(words)[k >> 1] = wd;
			}
			// best stop state (viterbi_detector.cc:342-350)
			const float m4 = __shfl_sync(0xffffffffu, pm, (h << 4) + 4), m12 = __shfl_sync(0xffffffffu, pm, (h << 4) + 12);
			__syncwarp();
			// ---- traceback (:371-391).  Serial part, one lane per burst: only the decision chain
			//      state' = (state >> 1) + 8 * decision, collected as a bit string G (bit k = decision at step k). ----
			unsigned gw[5] = { 0u, 0u, 0u, 0u, 0u };
			unsigned sF = (m12 > m4) ? 12u : 4u;
			if (s == 0) {
				unsigned state = sF;
#pragma unroll
				for (int wi = 4; wi >= 0; wi--) {
					unsigned acc = 0u;
					for (int k = min(N - 1, 32 * wi + 31); k >= 32 * wi; k--) {
						const unsigned g = (words[2 * k] >> ((h << 4) + state)) & 1u;
						acc |= g << (k & 31);
						state = (state >> 1) + (g << 3);
					}
					gw[wi] = acc;
				}
			}
			// ---- parallel part, lanes = steps.  Along the path state_k = g[k+1] g[k+2] g[k+3] g[k+4] (the start state
			//      supplies the bits beyond the end), which gives every step its parity-table bit; out_bit at step k is the
			//      XOR of (real_imag ^ parity) over the steps above k; the output sign is then one table lookup. ----
#pragma unroll
			for (int hh = 0; hh < 2; hh++) {
				const int bb = 2 * pair + hh;
				unsigned G[6];
#pragma unroll
				for (int wi = 0; wi < 5; wi++) G[wi] = __shfl_sync(0xffffffffu, gw[wi], hh << 4);
				G[5] = 0u;
				const unsigned sf = __shfl_sync(0xffffffffu, sF, hh << 4);
				// bits N .. N+3 = start state, most significant bit first
				const unsigned ext = ((sf >> 3) & 1u) | (((sf >> 2) & 1u) << 1) | (((sf >> 1) & 1u) << 2) | ((sf & 1u) << 3);
				if (N == 148) G[4] |= ext << 20; // bits N .. N+3 (N = 148 or 88: they stay inside one word)
				else G[2] |= ext << 24;
				const unsigned ri0 = ((N - 1) & 1) ? 0u : 1u;
				unsigned X[5], st_k[5];
#pragma unroll
				for (int wi = 0; wi < 5; wi++) {
					const int k = 32 * wi + lane;
					const unsigned nib = __funnelshift_rc(G[wi], G[wi + 1], lane + 1) & 15u; // g[k+1] .. g[k+4], LSB first
					st_k[wi] = __brev(nib) >> 28;
					const unsigned x = (ri0 ^ ((unsigned)(N - 1 - k) & 1u) ^ ((0x6666u >> st_k[wi]) & 1u)) & 1u;
					X[wi] = __ballot_sync(0xffffffffu, k < N && x);
				}
				unsigned carry = 0u; // parity of the X bits in the words above
#pragma unroll
				for (int wi = 4; wi >= 0; wi--) {
					const int k = 32 * wi + lane;
					const unsigned above = (lane == 31) ? 0u : (X[wi] >> (lane + 1));
					const unsigned out_bit = (__popc(above) + carry) & 1u;
					carry = (carry + __popc(X[wi])) & 1u;
					if (bb < p.n && k < N) {
						const unsigned g = (G[wi] >> lane) & 1u;
						const unsigned l = (words[2 * k + 1] >> ((hh << 4) + st_k[wi])) & 1u;
						const unsigned pos = (g != out_bit) ? l : g; // output[k] > 0
						p.bits[(size_t)bb * N + k] = pos ? (int8_t)-127 : (int8_t)127;
					}
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// First SCH acquisition: get_sch_buffer_chan_imp_resp (grgsm_vitac.cpp:298-309) over a capture of `len` samples
// (12 frames in ms_rx_lower.cpp:160-177).  get_chan_imp_resp's three loops (:183-235) become
//   sch_buffer_corr_kernel    one thread per search window: correlate_sequence over the 54 inner training symbols,
//                             4 samples apart (lanes = consecutive windows: coalesced), the reference's sequential
//                             float order, |c|^2 through double as std::pow(abs(c), 2) evaluates it;
//   sch_buffer_window_kernel  one CTA per capture: the 20-window energy is a RUNNING float sum in the reference
//                             (windowSum += p[i] - p[i-20]) and its first maximum decides the burst position, so the
//                             sum stays serial: the CTA stages 2,048 differences at a time in shared memory, thread 0
//                             adds them in order; then the 20 channel taps, corr_max and the start are written.
// detect_burst_nb at the position found is vitac_kernel with the taps just estimated (row_shift).
// ---------------------------------------------------------------------------------------------
struct SchBufParams {
	const float *bufs;
	int stride, offset, n, nwin;
	float2 *corr;   // [n][nwin]
	float *pw;	// [n][nwin]
	int32_t *start, *shift;
	int shift_lo, shift_hi;
	float *corr_max, *cir;
};

__global__ void __launch_bounds__(256)
sch_buffer_corr_kernel(SchBufParams p)
{
	const int b = blockIdx.y, w = blockIdx.x * 256 + threadIdx.x;
	if (w >= p.nwin) return;
	const float2 *x = reinterpret_cast<const float2 *>(p.bufs) + (size_t)b * p.stride + p.offset + w;
	const float2 *tseq = &c_tab.vitac_sch[5];
	float rr = 0.0f, ri = 0.0f;
#pragma unroll 6
	for (int ii = 0; ii < 54; ii++) {
		const float2 s = tseq[ii], v = __ldg(&x[ii * kOSR]);
		rr = fa(rr, fs(fm(s.x, v.x), fm(s.y, v.y)));
		ri = fa(ri, fa(fm(s.x, v.y), fm(s.y, v.x)));
	}
	const float2 c = make_float2(rr / 54.0f, -ri / 54.0f);
	p.corr[(size_t)b * p.nwin + w] = c;
	const float a = cabs_ref(c);
	p.pw[(size_t)b * p.nwin + w] = (float)((double)a * (double)a);
}

__global__ void __launch_bounds__(256)
sch_buffer_window_kernel(SchBufParams p)
{
	__shared__ float d[2048];
	__shared__ int s_best;
	const int b = blockIdx.x, tid = threadIdx.x;
	const float *pw = p.pw + (size_t)b * p.nwin;
	float ws = 0.0f, best = 0.0f;
	int bi = 0;
	if (tid == 0) {
		for (int i = 0; i < kCirLen; i++) ws = fa(ws, pw[i]);
		best = ws;
	}
	for (int i0 = kCirLen; i0 < p.nwin; i0 += 2048) {
		const int cnt = min(2048, p.nwin - i0);
		__syncthreads();
		for (int k = tid; k < cnt; k += 256) d[k] = fs(pw[i0 + k], pw[i0 + k - kCirLen]);
		__syncthreads();
		if (tid == 0) {
			int k = 0;
			for (; k + 8 <= cnt; k += 8) {
				float v[8];
#pragma unroll
				for (int u = 0; u < 8; u++) v[u] = d[k + u];
#pragma unroll
				for (int u = 0; u < 8; u++) {
					ws = fa(ws, v[u]);
					if (best < ws) { best = ws; bi = i0 + k + u - kCirLen + 1; }
				}
			}
			for (; k < cnt; k++) {
				ws = fa(ws, d[k]);
				if (best < ws) { best = ws; bi = i0 + k - kCirLen + 1; }
			}
		}
	}
	if (tid == 0) s_best = bi;
	__syncthreads();
	bi = s_best;
	if (tid < 32) {
		const float2 c = tid < kCirLen ? p.corr[(size_t)b * p.nwin + bi + tid] : make_float2(0.0f, 0.0f);
		float a = tid < kCirLen ? cabs_ref(c) : 0.0f;
#pragma unroll
		for (int o = 16; o; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
		if (tid < kCirLen) reinterpret_cast<float2 *>(p.cir)[(size_t)b * kCirLen + tid] = c;
		if (tid == 0) {
			const int st = bi - 47 * kOSR; // search_start_pos 0, search centre SYNC_POS + TRAIN_BEGINNING = 47 symbols
			p.corr_max[b] = a;
			p.start[b] = st;
			p.shift[b] = max(p.shift_lo, min(p.shift_hi, st));
		}
	}
}

} // namespace trxb200
