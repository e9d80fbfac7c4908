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

// shared memory per warp (floats): cb[2*nwin] | pw[nwin] | filt[2][160] | inc[2][8] | words[2][160] | misc[8]
__global__ void __launch_bounds__(128)
vitac_kernel(VitacParams p)
{
	extern __shared__ __align__(16) float vsm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	const int per_warp = (3 * p.nwin_max + 2 * 160 + 16 + 2 * 160 + 8 + 3) & ~3; // keep float2 alignment per warp
	float *base = vsm + (size_t)warp * per_warp;
	float2 *cb = reinterpret_cast<float2 *>(base);
	float *pw = base + 2 * p.nwin_max;
	float *filt = pw + p.nwin_max;	       // [2][160]
	float *inc = filt + 320;	       // [2][8]
	unsigned *words = reinterpret_cast<unsigned *>(inc + 16); // [160][2]: gt, lt
	int *misc = reinterpret_cast<int *>(words + 320);

	const int N = p.is_ab ? 88 : 148;
	const int center = p.is_ab ? 13 : 66;
	const int s0 = (center - 5) * kOSR + 1;
	const int s1 = (center + 5 + 5 + (p.is_ab ? p.max_delay : 0)) * kOSR;
	const int nwin = s1 - s0;
	const int tlen = p.is_ab ? 31 : 16;

	const int npairs = (p.n + 1) >> 1;
	for (int pair = blockIdx.x * wpb + warp; pair < npairs; pair += gridDim.x * wpb) {
		__syncwarp();
		for (int h = 0; h < 2; h++) {
			const int b = 2 * pair + h;
			if (b >= p.n) {
				for (int i = lane; i < 160; i += 32) filt[h * 160 + i] = 0.0f;
				if (lane < 8) inc[h * 8 + lane] = 0.0f;
				continue;
			}
			const float2 *in = reinterpret_cast<const float2 *>(p.bufs) + (size_t)b * p.stride + p.offset;
			const float2 *tseq = p.is_ab ? &c_tab.vitac_access[5] : &c_tab.vitac_norm[p.tsc[b] > 8 ? 8 : p.tsc[b]][5];
			// ---- correlation per search window (correlate_sequence :148-156) ----
			for (int w = lane; w < nwin; w += 32) {
				const float2 *x = in + s0 + w;
				float rr = 0.0f, ri = 0.0f;
				for (int ii = 0; ii < tlen; ii++) {
					const float2 s = tseq[ii], v = __ldg(&x[ii * kOSR]);
					rr = fa(rr, fs(fm(s.x, v.x), fm(s.y, v.y)));
					ri = fa(ri, fa(fm(s.x, v.y), fm(s.y, v.x)));
				}
				const float2 c = make_float2(rr / (float)tlen, -ri / (float)tlen);
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
			const int bi = misc[h];
			int st = s0 + bi - center * kOSR;
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
			// ---- matched filter (mafi :168-181), only the component the trellis reads ----
			for (int nn = lane; nn < N; nn += 32) {
				const float2 *x = in + st + nn * kOSR;
				float acc = 0.0f;
				const bool want_imag = !(nn & 1);
				for (int ii = 0; ii < kCirLen; ii++) {
					if (nn * kOSR + ii >= N * kOSR) break;
					const float2 v = __ldg(&x[ii]), c = cir[ii];
					const float t = want_imag ? fa(fm(v.x, c.y), fm(v.y, c.x)) : fs(fm(v.x, c.x), fm(v.y, c.y));
					acc = fa(acc, t);
				}
				filt[h * 160 + nn] = acc;
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
			float pm = (s == 3) ? 0.0f : (float)(-10e30);
			const int src1 = (h << 4) + pp, src2 = src1 + 8;
			const float *f = filt + h * 160;
			for (int k = 0; k < N; k++) {
				const float o1 = __shfl_sync(0xffffffffu, pm, src1), o2 = __shfl_sync(0xffffffffu, pm, src2);
				const float x = f[k];
				float c1, c2;
				if (!(k & 1)) { // imaginary step
					const float sx = odd ? -x : x;
					c1 = fa(fa(o1, sx), i1I);
					c2 = fa(fa(o2, sx), i2I);
				} else {
					const float sx = odd ? x : -x;
					c1 = fa(fa(o1, sx), i1R);
					c2 = fa(fa(o2, sx), i2R);
				}
				const float d = fs(c2, c1);
				pm = (d < 0.0f) ? c1 : c2;
				const unsigned gt = __ballot_sync(0xffffffffu, d > 0.0f), lt = __ballot_sync(0xffffffffu, d < 0.0f);
				if (lane == 0) { words[2 * k] = gt; words[2 * k + 1] = lt; }
			}
			// best stop state (viterbi_detector.cc:342-350)
			const float m4 = __shfl_sync(0xffffffffu, pm, (h << 4) + 4), m12 = __shfl_sync(0xffffffffu, pm, (h << 4) + 12);
			__syncwarp();
			// ---- traceback (:371-391), one lane per burst ----
			if (s == 0 && 2 * pair + h < p.n) {
				unsigned state = (m12 > m4) ? 12u : 4u;
				unsigned out_bit = 0, real_imag = ((N - 1) & 1) ? 0u : 1u;
				int8_t *ob = p.bits + (size_t)(2 * pair + h) * N;
				const unsigned par = 0x6666u; // parity_table bits
				for (int k = N - 1; k >= 0; k--) {
					const unsigned sh = (h << 4) + state;
					const unsigned g = (words[2 * k] >> sh) & 1u, l = (words[2 * k + 1] >> sh) & 1u;
					const unsigned decision = g;
					const unsigned pos = (decision != out_bit) ? l : g; // output[k] > 0
					ob[k] = pos ? (int8_t)-127 : (int8_t)127;
					out_bit = out_bit ^ real_imag ^ ((par >> state) & 1u);
					state = (state >> 1) + (decision ? 8u : 0u);
					real_imag ^= 1u;
				}
			}
		}
	}
}

} // namespace trxb200
