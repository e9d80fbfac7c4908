// modulate.cu — batched GMSK (Laurent C0+C1) and 8-PSK modulators for sm_100a.
//
// Reference: modulateBurstLaurent (sigProcLib.cpp:595-670) and shapeEdgeBurst (:739-763).  Both
// build a 625-sample vector that is zero except at every 4th sample and push it through the 16-tap
// C0 (and 8-tap C1) pulse with sse_conv_real16/8.  Because only taps k == (15-n) mod 4 meet a
// non-zero sample, exactly one SSE lane carries data and the reference's summation tree collapses
// to (p_a + p_b) + (p_c + p_d) for C0 and p_a + p_b for C1 — four (two) products per output,
// evaluated here in that order, so the waveform is bit-identical (up to the sign of exact zeros).
// One thread per output sample, one warp-group of 160 threads... simply: block = 4 bursts x 160
// threads would waste lanes, so a block of 256 threads walks its bursts' samples linearly and writes
// fully coalesced float2.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

__global__ void __launch_bounds__(256)
modulate_gmsk_kernel(const uint8_t *__restrict__ bits, int nbits, int bits_stride, int n, float *__restrict__ out,
		     int out_stride, const float *__restrict__ mtab)
{
	const float2 *__restrict__ rot4 = reinterpret_cast<const float2 *>(mtab + kModRot4);
	const float *__restrict__ pc0 = mtab + kModC0, *__restrict__ pc1 = mtab + kModC1;
	__shared__ float sym[2][160];  // NRZ symbols incl. the two padded "0" symbols
	__shared__ float ph1[2][160];  // C1 phase sign per symbol slot
	for (int b0 = blockIdx.x * 2; b0 < n; b0 += gridDim.x * 2) {
		__syncthreads();
		for (int t = threadIdx.x; t < 2 * 160; t += blockDim.x) {
			const int w = t / 160, m = t % 160, b = b0 + w;
			float s = 0.0f, ph = 0.0f;
			if (b < n && m <= nbits + 1 && 4 * m < 625) {
				const uint8_t *bb = bits + (size_t)b * bits_stride;
				s = (m == 0 || m == nbits + 1) ? -1.0f : (float)(2.0 * (bb[m - 1] & 1) - 1.0);
				if (m == 2) ph = -1.0f;
				else if (m >= 3) ph = (float)(2.0 * ((bb[m - 2] & 1) ^ (bb[m - 3] & 1)) - 1.0);
			}
			sym[w][m] = s;
			ph1[w][m] = ph;
		}
		__syncthreads();
		for (int t = threadIdx.x; t < 2 * 625; t += blockDim.x) {
			const int w = t / 625, nn = t % 625, b = b0 + w;
			if (b >= n) break;
			const int js = (15 - nn) & 3;
			// C0: samples 4m = nn-15+k, k = js + 4g
			float pr[4], pi[4];
#pragma unroll
			for (int g = 0; g < 4; g++) {
				const int k = js + 4 * g, idx = nn - 15 + k, m = idx >> 2;
				float xr = 0.0f, xi = 0.0f;
				if (idx >= 0 && m < 160) {
					const float2 r = __ldg(&rot4[idx]);
					xr = fm(r.x, sym[w][m]);
					xi = fm(r.y, sym[w][m]);
				}
				pr[g] = fm(xr, __ldg(&pc0[k]));
				pi[g] = fm(xi, __ldg(&pc0[k]));
			}
			float yr = fa(fa(pr[0], pr[1]), fa(pr[2], pr[3]));
			float yi = fa(fa(pi[0], pi[1]), fa(pi[2], pi[3]));
			// C1: samples 4m = nn-7+k, k = js + 4g, c1 = c0 * (0, ph)
			float qr[2], qi[2];
#pragma unroll
			for (int g = 0; g < 2; g++) {
				const int k = js + 4 * g, idx = nn - 7 + k, m = idx >> 2;
				float xr = 0.0f, xi = 0.0f;
				if (idx >= 0 && m < 160) {
					const float2 r = __ldg(&rot4[idx]);
					const float c0r = fm(r.x, sym[w][m]), c0i = fm(r.y, sym[w][m]);
					const float ph = ph1[w][m];
					xr = fs(fm(c0r, 0.0f), fm(c0i, ph));
					xi = fa(fm(c0r, ph), fm(c0i, 0.0f));
				}
				qr[g] = fm(xr, __ldg(&pc1[k]));
				qi[g] = fm(xi, __ldg(&pc1[k]));
			}
			yr = fa(yr, fa(qr[0], qr[1]));
			yi = fa(yi, fa(qi[0], qi[1]));
			reinterpret_cast<float2 *>(out)[(size_t)b * out_stride + nn] = make_float2(yr, yi);
		}
	}
}

__global__ void __launch_bounds__(256)
modulate_edge_kernel(const uint8_t *__restrict__ bits, int nbits, int bits_stride, int n, float *__restrict__ out,
		     int out_stride, const float *__restrict__ mtab)
{
	const float *__restrict__ pc0 = mtab + kModC0;
	const float2 *__restrict__ erot = reinterpret_cast<const float2 *>(mtab + kModEdgeRot);
	const float2 *__restrict__ psk8 = reinterpret_cast<const float2 *>(mtab + kModPsk8);
	__shared__ float2 sym[2][160]; // rotated symbols at sample 4 + 4i -> slot m = i + 1
	int nsym = nbits / 3;
	if (nsym * 4 > 625) nsym = 156;
	for (int b0 = blockIdx.x * 2; b0 < n; b0 += gridDim.x * 2) {
		__syncthreads();
		for (int t = threadIdx.x; t < 2 * 160; t += blockDim.x) {
			const int w = t / 160, m = t % 160, b = b0 + w;
			float2 v = make_float2(0.0f, 0.0f);
			if (b < n && m >= 1 && m <= nsym && 4 * m < 625) {
				const uint8_t *bb = bits + (size_t)b * bits_stride + 3 * (m - 1);
				const unsigned idx = (bb[0] & 1u) | ((bb[1] & 1u) << 1) | ((bb[2] & 1u) << 2);
				v = cmul_exact(__ldg(&psk8[idx]), __ldg(&erot[m - 1]));
			}
			sym[w][m] = v;
		}
		__syncthreads();
		for (int t = threadIdx.x; t < 2 * 625; t += blockDim.x) {
			const int w = t / 625, nn = t % 625, b = b0 + w;
			if (b >= n) break;
			const int js = (15 - nn) & 3;
			float pr[4], pi[4];
#pragma unroll
			for (int g = 0; g < 4; g++) {
				const int k = js + 4 * g, idx = nn - 15 + k, m = idx >> 2;
				float2 xv = make_float2(0.0f, 0.0f);
				if (idx >= 0 && m < 160) xv = sym[w][m];
				pr[g] = fm(xv.x, __ldg(&pc0[k]));
				pi[g] = fm(xv.y, __ldg(&pc0[k]));
			}
			reinterpret_cast<float2 *>(out)[(size_t)b * out_stride + nn] =
				make_float2(fa(fa(pr[0], pr[1]), fa(pr[2], pr[3])), fa(fa(pi[0], pi[1]), fa(pi[2], pi[3])));
		}
	}
}

} // namespace trxb200
