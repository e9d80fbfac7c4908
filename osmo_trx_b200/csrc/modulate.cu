// modulate.cu — batched GMSK (Laurent C0+C1) and 8-PSK modulators for sm_100a.
//
// Reference: modulateBurstLaurent (sigProcLib.cpp:595-670) and shapeEdgeBurst (:739-763).  Both
// build a 625-sample vector that is zero except at every 4th sample and push it through the 16-tap
// C0 (and 8-tap C1) pulse with sse_conv_real16/8.  Because only taps k == (15-n) mod 4 meet a
// non-zero sample, exactly one SSE lane carries data and the reference's summation tree collapses
// to (p_a + p_b) + (p_c + p_d) for C0 and p_a + p_b for C1 — four (two) products per output,
// evaluated here in that order, so the waveform is bit-identical (up to the sign of exact zeros).
//
// Work mapping (HBM-write bound: 148 B in, 5,000 B out per burst): one warp per burst.
//   symbols  the burst's bits are read once (one byte per lane) and turned into five ballot words; the
//            three bits each symbol slot needs come out of funnel shifts of those words, so the rotated
//            symbols x[m] = rot4[4m] * nrz[m] and the C1 stream c1[m] are built with no further loads
//            and written to a per-warp shared row (8-PSK: three byte loads per symbol, Gray map, rotation).
//   outputs  output n = 4q + r only meets symbols m = q-3..q (C0) and q-1..q (C1) with taps 3-r+4g; lane
//            = n mod 32 keeps r, hence its six taps, in registers for the whole kernel, reads the symbols
//            with broadcast LDS.64 and evaluates the (re,im) pairs on the packed FP32 pipe (exact FMUL2 /
//            FADD2, no contraction).  Each warp store instruction writes 256 contiguous bytes.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

constexpr int kModWarps = 8;
constexpr int kModRow = 168; // symbol slots m = -3 .. 164 of one burst (index m + 3)

// bit `lane + 32*i - back` of the 160-bit ballot bitmap w[0..4] (zero outside), i compile-time
__device__ __forceinline__ unsigned mod_bit(const unsigned (&w)[5], int i, int back, int lane)
{
	const unsigned lo = (i >= 1 && i <= 5) ? w[i - 1] : 0u;
	const unsigned hi = (i >= 0 && i <= 4) ? w[i] : 0u;
	return (__funnelshift_r(lo, hi, 32 - back) >> lane) & 1u;
}

__global__ void __launch_bounds__(kModWarps * 32)
modulate_gmsk_kernel(const uint8_t *__restrict__ bits, int nbits, int bits_stride, int n, float *__restrict__ out,
		     int out_stride, const float *__restrict__ mtab, float negzero)
{
	const float2 *__restrict__ rot4 = reinterpret_cast<const float2 *>(mtab + kModRot4);
	const float *__restrict__ pc0 = mtab + kModC0, *__restrict__ pc1 = mtab + kModC1;
	__shared__ float2 xs[kModWarps][kModRow];  // C0 stream: rotated NRZ symbols incl. the two padded "0" symbols
	__shared__ float2 c1s[kModWarps][kModRow]; // C1 stream
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const float2 nz = make_float2(negzero, negzero);
	const int r = lane & 3;
	float t0[4], t1[2];
#pragma unroll
	for (int g = 0; g < 4; g++) t0[g] = __ldg(&pc0[3 - r + 4 * g]);
#pragma unroll
	for (int g = 0; g < 2; g++) t1[g] = __ldg(&pc1[3 - r + 4 * g]);
	// the rotator at the symbol instants this lane builds (slot j = lane + 32*i, m = j - 3)
	float2 rt[6];
#pragma unroll
	for (int i = 0; i < 6; i++) {
		const int m = lane + 32 * i - 3;
		rt[i] = (m >= 0 && 4 * m < 625) ? __ldg(&rot4[4 * m]) : make_float2(0.0f, 0.0f);
	}
	float2 *xw = xs[warp], *cw = c1s[warp];
	for (int b = blockIdx.x * kModWarps + warp; b < n; b += gridDim.x * kModWarps) {
		const uint8_t *bb = bits + (size_t)b * bits_stride;
		unsigned w[5];
#pragma unroll
		for (int i = 0; i < 5; i++) {
			const int k = lane + 32 * i;
			w[i] = __ballot_sync(0xffffffffu, k < nbits && (bb[k] & 1));
		}
		__syncwarp(); // the previous burst's output pass is done with the rows
#pragma unroll
		for (int i = 0; i < 6; i++) {
			const int j = lane + 32 * i, m = j - 3;
			if (j < kModRow) {
				// bits m-1, m-2, m-3  <->  bitmap index j - 4, j - 5, j - 6
				const unsigned b1 = mod_bit(w, i, 4, lane), b2 = mod_bit(w, i, 5, lane), b3 = mod_bit(w, i, 6, lane);
				float s = 0.0f, ph = 0.0f;
				if (m >= 0 && m <= nbits + 1 && 4 * m < 625) {
					s = (m == 0 || m == nbits + 1) ? -1.0f : (b1 ? 1.0f : -1.0f);
					if (m == 2) ph = -1.0f;
					else if (m >= 3) ph = (b2 ^ b3) ? 1.0f : -1.0f;
				}
				const float c0r = fm(rt[i].x, s), c0i = fm(rt[i].y, s);
				xw[j] = make_float2(c0r, c0i);
				cw[j] = make_float2(fs(fm(c0r, 0.0f), fm(c0i, ph)), fa(fm(c0r, ph), fm(c0i, 0.0f)));
			}
		}
		__syncwarp();
		float2 *orow = reinterpret_cast<float2 *>(out) + (size_t)b * out_stride;
#pragma unroll 4
		for (int i = 0; i < 20; i++) {
			const int nn = lane + 32 * i;
			if (nn < 625) {
				const int q = nn >> 2; // symbols m = q-3+g  <->  row index q+g
				const float2 p0 = mul2(xw[q], bc2(t0[0]), nz), p1 = mul2(xw[q + 1], bc2(t0[1]), nz);
				const float2 p2 = mul2(xw[q + 2], bc2(t0[2]), nz), p3 = mul2(xw[q + 3], bc2(t0[3]), nz);
				const float2 q0 = mul2(cw[q + 2], bc2(t1[0]), nz), q1 = mul2(cw[q + 3], bc2(t1[1]), nz);
				orow[nn] = add2(add2(add2(p0, p1), add2(p2, p3)), add2(q0, q1));
			}
		}
	}
}

__global__ void __launch_bounds__(kModWarps * 32)
modulate_edge_kernel(const uint8_t *__restrict__ bits, int nbits, int bits_stride, int n, float *__restrict__ out,
		     int out_stride, const float *__restrict__ mtab, float negzero)
{
	const float *__restrict__ pc0 = mtab + kModC0;
	const float2 *__restrict__ erot = reinterpret_cast<const float2 *>(mtab + kModEdgeRot);
	const float2 *__restrict__ psk8 = reinterpret_cast<const float2 *>(mtab + kModPsk8);
	__shared__ float2 xs[kModWarps][kModRow]; // rotated symbols at sample 4 + 4i -> slot m = i + 1
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const float2 nz = make_float2(negzero, negzero);
	const int r = lane & 3;
	float t0[4];
#pragma unroll
	for (int g = 0; g < 4; g++) t0[g] = __ldg(&pc0[3 - r + 4 * g]);
	float2 rt[6];
#pragma unroll
	for (int i = 0; i < 6; i++) {
		const int m = lane + 32 * i - 3;
		rt[i] = (m >= 1 && m <= 156) ? __ldg(&erot[m - 1]) : make_float2(0.0f, 0.0f);
	}
	int nsym = nbits / 3;
	if (nsym * 4 > 625) nsym = 156;
	float2 *xw = xs[warp];
	for (int b = blockIdx.x * kModWarps + warp; b < n; b += gridDim.x * kModWarps) {
		const uint8_t *bb = bits + (size_t)b * bits_stride;
		__syncwarp();
#pragma unroll
		for (int i = 0; i < 6; i++) {
			const int j = lane + 32 * i, m = j - 3;
			if (j < kModRow) {
				float2 v = make_float2(0.0f, 0.0f);
				if (m >= 1 && m <= nsym && 4 * m < 625) {
					const uint8_t *b3 = bb + 3 * (m - 1);
					const unsigned idx = (b3[0] & 1u) | ((b3[1] & 1u) << 1) | ((b3[2] & 1u) << 2);
					v = cmul_exact(__ldg(&psk8[idx]), rt[i]);
				}
				xw[j] = v;
			}
		}
		__syncwarp();
		float2 *orow = reinterpret_cast<float2 *>(out) + (size_t)b * out_stride;
#pragma unroll 4
		for (int i = 0; i < 20; i++) {
			const int nn = lane + 32 * i;
			if (nn < 625) {
				const int q = nn >> 2;
				const float2 p0 = mul2(xw[q], bc2(t0[0]), nz), p1 = mul2(xw[q + 1], bc2(t0[1]), nz);
				const float2 p2 = mul2(xw[q + 2], bc2(t0[2]), nz), p3 = mul2(xw[q + 3], bc2(t0[3]), nz);
				orow[nn] = add2(add2(p0, p1), add2(p2, p3));
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// modulate_basic_kernel — the modulator forms the reference uses outside the 4-sps transmit path: building its
// correlation references at setup (sigProcLib.cpp:1227-1465) and transmitting at 1 sps.  One warp per burst, lanes =
// output samples; exact float32 order of the reference (no contraction).
//   mode 0  modulateBurstBasic, sps = 1 (:938-967): NRZ bits x GMSKRotation1 (complex x real), then the 4-tap GSMPulse1
//           through sse_conv_real4 (START_ONLY: y[i] = sum_k x[i-3+k] h[k], zero head-room; (L0+L1)+(L2+L3), L_k = one product)
//   mode 1  rotateBurst (:558-580): NRZ impulses at stride sps x the rotator (complex x complex), through the one-tap
//           "empty" pulse (base_convolve_real: 0 + x * 1)
//   mode 2  rotateEdgeBurst (:672-689): Gray-mapped 8-PSK symbols x e^{j i 3 pi / 8} at stride sps, zeros between
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
modulate_basic_kernel(const uint8_t *__restrict__ bits, int nbits, int bits_stride, int n, int guard, int sps, int mode,
		      float *__restrict__ out, int out_stride)
{
	const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	const int nsym = mode == 2 ? nbits / 3 : nbits;
	const int olen = sps * (nsym + guard);
	for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
		const uint8_t *bb = bits + (size_t)b * bits_stride;
		float2 *o = reinterpret_cast<float2 *>(out) + (size_t)b * out_stride;
		for (int i = lane; i < olen; i += 32) {
			float2 y = make_float2(0.0f, 0.0f);
			if (mode == 0) {
				float lr[4], li[4];
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const int m = i - 3 + k;
					float2 x = make_float2(0.0f, 0.0f);
					if (m >= 0 && m < nbits) {
						const float sgn = (float)(2.0 * (double)(bb[m] & 1) - 1.0);
						x = make_float2(fm(c_tab.rot1[m].x, sgn), fm(c_tab.rot1[m].y, sgn));
					}
					lr[k] = fm(x.x, c_tab.pulse1_c0[k]);
					li[k] = fm(x.y, c_tab.pulse1_c0[k]);
				}
				y = make_float2(fa(fa(lr[0], lr[1]), fa(lr[2], lr[3])), fa(fa(li[0], li[1]), fa(li[2], li[3])));
			} else if (i % sps == 0 && i / sps < nsym) {
				const int m = i / sps;
				if (mode == 1) {
					const float2 sgn = make_float2((float)(2.0 * (double)(bb[m] & 1) - 1.0), 0.0f);
					const float2 x = cmul_exact(sps == 1 ? c_tab.rot1[i] : c_tab.rot4[i], sgn);
					y = make_float2(fa(0.0f, fm(x.x, 1.0f)), fa(0.0f, fm(x.y, 1.0f)));
				} else {
					const unsigned idx = (bb[3 * m] & 1u) | ((bb[3 * m + 1] & 1u) << 1) | ((bb[3 * m + 2] & 1u) << 2);
					y = cmul_exact(c_tab.psk8[idx], c_tab.edge_mod_rot[m]);
				}
			}
			o[i] = y;
		}
	}
}

} // namespace trxb200
