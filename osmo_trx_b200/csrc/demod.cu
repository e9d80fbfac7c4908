// demod.cu — batched soft demodulation (demodAnyBurst, sigProcLib.cpp:2130-2137) for sm_100a.
//
// Reference chain per burst (demodCommon :2030-2048 + demodGmskBurst :2055-2072):
//   delayVector(-toa*4): 20-tap fractional-delay FIR over all 625 samples + integer shift
//   scaleVector(1/amp), downsampleBurst: 16-tap decimating FIR (only 156 outputs are kept)
//   GMSKReverseRotate at 1 sps, real part.
// All of it is linear and only every 4th sample of the delayed burst survives, so the kernel
// evaluates the surviving outputs directly with the 35-tap composite filter delay[f] (*) decimator
// built on the host (tables.cpp, `comp`): 3x fewer MACs than the two-stage form, one HBM read per
// burst.  Where the reference's intermediate vectors are truncated (samples shifted in from outside
// the 625-sample burst are zero, the decimator's history before sample 0 is zero) the affected
// leading outputs use the composite truncated to the surviving decimator taps (comp[f][kmin]); outputs
// truncated from above take a generic two-stage path.  The complex gain 1/amp is applied while the
// burst is staged into shared memory, and because e^{-j*pi*n/2} only selects +-re/+-im (the table
// entries are exactly +-1 on the selected component; the other has ~1e-16 leakage from the
// double-precision phase accumulation, far below the 1e-4 soft-bit tolerance) a GMSK output needs
// one real 35-tap dot product.  This chain feeds no decisions, so FMA is used.
//
// Work mapping, one warp per burst:
//   stage   16-byte global loads (two samples) of the burst, scaled by 1/amp, into a polyphase-planar
//           window u[comp][phase][q] (window sample 4q+phase); the window origin is chosen 4-aligned
//           to the 16-byte grid of the row, the residual shift e is folded into the tap index.
//   FIR     lane l owns 5 consecutive outputs 5l..5l+4: for each polyphase plane it loads the 13 (+11)
//           window values its outputs share and the plane's 9 taps (shared-memory broadcast), and
//           issues 45 FMAs into 5 independent accumulators.  Lane stride 5 words = conflict free.
//   edges   the few leading outputs whose decimator taps are truncated are recomputed with the
//           truncated composites, 4 lanes per output + shuffle reduction.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {
namespace {

constexpr int kPlane = 168;		 // floats per polyphase plane (q = 0..165 used)
constexpr int kCompWords = 4 * kPlane;	 // one component
constexpr int kNQ = 166;		 // staged q range: window samples 0 .. 663
constexpr int kCoefWords = 40;		 // shifted composite taps ce[0..39]
constexpr int kDemodWarpFloats = 2 * kCompWords + kCoefWords + 2 * 164; // + coef + complex dec scratch (EDGE)

__device__ __forceinline__ float win_get(const float *u, int comp, int w)
{
	return u[comp * kCompWords + (w & 3) * kPlane + (w >> 2)];
}

// generic two-stage evaluation of one decimated sample restricted to decimator taps [kmin,kmax]
// (only when the shifted burst runs off the top of the 625-sample vector); w0 = window index of tap 0
__device__ __noinline__ float2 slow_output(const float *u, int w0, int f, int kmin, int kmax)
{
	float2 acc = make_float2(0.0f, 0.0f);
	for (int k = kmin; k <= kmax; k++) {
		float yr = 0.0f, yi = 0.0f;
		if (f < 64) {
			for (int j = 0; j < 20; j++) {
				const float h = c_tab.delay[f][j];
				yr = fmaf(win_get(u, 0, w0 + k + j), h, yr);
				yi = fmaf(win_get(u, 1, w0 + k + j), h, yi);
			}
		} else {
			yr = win_get(u, 0, w0 + k + 9);
			yi = win_get(u, 1, w0 + k + 9);
		}
		acc.x = fmaf(yr, c_tab.dnsamp[k], acc.x);
		acc.y = fmaf(yi, c_tab.dnsamp[k], acc.y);
	}
	return acc;
}

// Re(e^{-j*pi*i/2} * d): GMSKReverseRotate(1 sps) + real part (sigProcLib.cpp:262-287,2011-2022)
__device__ __forceinline__ float derot_real(int i, float dr, float di)
{
	const float v = (i & 1) ? di : dr;
	return (i & 2) ? -v : v;
}

// ---- EDGE: demodEdgeBurst :2105-2128 on the staged window (complex outputs) ----
__device__ __noinline__ void demod_edge_burst(const DemodParams &p, int b, float *u, float2 *decs, const float *ce, int e, int f,
					      int whole, int lane)
{
	for (int i = lane; i < 160; i += 32) {
		float2 d = make_float2(0.0f, 0.0f);
		if (i < 156) {
			const int kmin = max(0, max(15 - 4 * i, 15 - 4 * i + whole));
			const int kmax = min(15, 639 + whole - 4 * i);
			if (kmin <= kmax) {
				if (kmax == 15) {
					const float *__restrict__ c = p.comp + ((size_t)f * 16 + kmin) * 36;
#pragma unroll 5
					for (int t = 0; t < 35; t++) {
						const float ct = __ldg(&c[t]);
						const int w = 4 * i + t + e;
						const int q = (w & 3) * kPlane + (w >> 2);
						d.x = fmaf(u[q], ct, d.x);
						d.y = fmaf(u[kCompWords + q], ct, d.y);
					}
				} else {
					d = slow_output(u, 4 * i + e, f, kmin, kmax);
				}
			}
			decs[2 + i] = d;
		}
	}
	if (lane < 2) { decs[lane] = make_float2(0.0f, 0.0f); decs[158 + lane] = make_float2(0.0f, 0.0f); }
	__syncwarp();
	float err = 0.0f;
	for (int i = lane; i < 160; i += 32) {
		float2 rot = make_float2(0.0f, 0.0f);
		if (i < 156) {
			// 5-tap static equaliser, NO_DELAY span, sequential MAC (convolve_base.c:27-60)
			float er = 0.0f, ei = 0.0f;
#pragma unroll
			for (int k = 0; k < 5; k++) {
				er = fa(er, fm(decs[i + k].x, c_tab.c0_inv[k]));
				ei = fa(ei, fm(decs[i + k].y, c_tab.c0_inv[k]));
			}
			rot = cmul_exact(make_float2(er, ei), p.edge_tab[i & 15]); // derotateEdgeBurst :691-711
			if (i >= 8 && i < 148) {
				// computeEdgeCI :2074-2093
				const float step = 2.0f * 3.14159274f / 8.0f;
				int k = (int)roundf(atan2f(rot.y, rot.x) / step);
				k = min(max(k, -4), 4);
				const float2 ideal = p.edge_tab[16 + k + 4];
				const float2 er2 = make_float2(fs(ideal.x, rot.x), fs(ideal.y, rot.y));
				err += norm2(er2);
			}
		}
		// softSliceEdgeBurst :1962-2006
		if (i < 148) {
			const float2 r1 = cmul_exact(rot, c_tab.edge_rot1);
			float *o = p.soft + (size_t)b * p.soft_stride + 3 * i;
			o[0] = -r1.y;
			o[1] = r1.x;
			const float2 r2 = cmul_exact(make_float2(fabsf(r1.x), fabsf(r1.y)), c_tab.edge_rot2);
			o[2] = -r2.y;
		}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1)
		err += __shfl_xor_sync(0xffffffffu, err, o);
	if (lane == 0)
		p.ci[b] = fm(3.0103f, log2f(140.0f / err));
}

} // namespace

__global__ void __launch_bounds__(256, 3)
demod_kernel(DemodParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	float *u = reinterpret_cast<float *>(smem_raw) + (size_t)warp * kDemodWarpFloats;
	float *ce = u + 2 * kCompWords;
	float2 *decs = reinterpret_cast<float2 *>(ce + kCoefWords);
	const unsigned base_par = (unsigned)((reinterpret_cast<uintptr_t>(p.bursts) >> 3) & 1u);

	for (int b = blockIdx.x * wpb + warp; b < p.n; b += gridDim.x * wpb) {
		const int rc = p.rc[b];
		const float2 *x = reinterpret_cast<const float2 *>(p.bursts) + (size_t)b * p.stride;

		if (rc <= 0) {
			// undetected burst: only the deferred clipping report is left to do (sigProcLib.cpp:1746-1764)
			if (p.fix_clip && rc == 0) {
				float mx = 0.0f;
				for (int i = lane; i < 625; i += 32) {
					const float2 v = __ldg(&x[i]);
					mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
				}
#pragma unroll
				for (int o = 16; o; o >>= 1)
					mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
				if (lane == 0 && mx > 30000.0f) {
					p.rc[b] = -2;
					if (p.flags) p.flags[b] |= 4;
				}
			}
			continue;
		}

		// ---- per-burst scalars (demodCommon / delayVector :1046-1060) ----
		const float2 amp = reinterpret_cast<const float2 *>(p.amp)[b];
		const float an = norm2(amp);
		const float2 s = make_float2(amp.x / an, -amp.y / an); // (complex)1.0 / amp
		const float delay = fm(-p.toa[b], 4.0f);
		const int whole = (int)floorf(delay);
		const float frac = fs(delay, (float)whole);
		int f = 64;
		if ((double)fabsf(frac) > 1e-2) {
			f = (int)floorf(fm(frac, 64.0f));
			f = min(max(f, 0), 63);
		}
		const int off = -24 - whole;		   // window sample 0 <-> burst sample `off` (before alignment)
		const unsigned row_par = (base_par + (unsigned)(((size_t)b * (size_t)p.stride) & 1u)) & 1u;
		const int e = (int)(((unsigned)off - row_par) & 1u); // residual shift so that off2 has the row's 16B parity
		const int off2 = off - e;		   // window sample w <-> burst sample w + off2; output i tap t reads w = 4i+t+e
		const bool edge = (rc == 5);

		__syncwarp();
		// shifted composite taps: ce[u] = comp[f][0][u - e]
		{
			const float *__restrict__ c = p.comp + (size_t)f * 16 * 36;
			for (int k = lane; k < kCoefWords; k += 32) {
				const int t = k - e;
				ce[k] = (t >= 0 && t < 35) ? __ldg(&c[t]) : 0.0f;
			}
		}
		// ---- stage s*x into the planar window, zero outside the burst; clip scan rides along ----
		// All 16-byte loads of the burst are issued before the first use (12 in flight per lane) so the warp
		// pays the HBM latency once per burst; groups straddling the burst ends are patched afterwards.
		float mx = 0.0f;
		{
			float4 ld[6][2];
#pragma unroll
			for (int it = 0; it < 6; it++) {
				const int q = lane + 32 * it;
				const int p0 = off2 + 4 * q;
				const bool full = (q < kNQ) && p0 >= 0 && p0 <= 621;
				ld[it][0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				ld[it][1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				if (full) {
					ld[it][0] = __ldg(reinterpret_cast<const float4 *>(x + p0));
					ld[it][1] = __ldg(reinterpret_cast<const float4 *>(x + p0 + 2));
				}
			}
#pragma unroll
			for (int it = 0; it < 6; it++) {
				const int q = lane + 32 * it;
				if (q < kNQ) {
					const float4 a = ld[it][0], c = ld[it][1];
					mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
					mx = fmaxf(mx, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
					float *ur = u + q, *ui = u + kCompWords + q;
					ur[0 * kPlane] = fmaf(a.x, s.x, -a.y * s.y); ui[0 * kPlane] = fmaf(a.x, s.y, a.y * s.x);
					ur[1 * kPlane] = fmaf(a.z, s.x, -a.w * s.y); ui[1 * kPlane] = fmaf(a.z, s.y, a.w * s.x);
					ur[2 * kPlane] = fmaf(c.x, s.x, -c.y * s.y); ui[2 * kPlane] = fmaf(c.x, s.y, c.y * s.x);
					ur[3 * kPlane] = fmaf(c.z, s.x, -c.w * s.y); ui[3 * kPlane] = fmaf(c.z, s.y, c.w * s.x);
				}
			}
			// partial groups at the two ends of the burst (at most two lanes per burst have one)
#pragma unroll
			for (int it = 0; it < 6; it++) {
				const int q = lane + 32 * it;
				const int p0 = off2 + 4 * q;
				if (q < kNQ && ((p0 < 0 && p0 > -4) || (p0 > 621 && p0 < 625))) {
					for (int j = 0; j < 4; j++) {
						const int src = p0 + j;
						float2 v = make_float2(0.0f, 0.0f);
						if (src >= 0 && src < 625) v = __ldg(&x[src]);
						mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
						u[j * kPlane + q] = fmaf(v.x, s.x, -v.y * s.y);
						u[kCompWords + j * kPlane + q] = fmaf(v.x, s.y, v.y * s.x);
					}
				}
			}
		}
		if (p.flags && p.fix_clip) {
			// samples the window did not cover (only for extreme shifts)
			if (off2 > 0 || off2 + 4 * kNQ < 625) {
				for (int i = lane; i < 625; i += 32) {
					if (i - off2 < 0 || i - off2 >= 4 * kNQ) {
						const float2 raw = __ldg(&x[i]);
						mx = fmaxf(mx, fmaxf(fabsf(raw.x), fabsf(raw.y)));
					}
				}
			}
#pragma unroll
			for (int o = 16; o; o >>= 1)
				mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
			if (lane == 0 && mx > 30000.0f) p.flags[b] |= 4;
		}
		__syncwarp();

		if (edge) {
			demod_edge_burst(p, b, u, decs, ce, e, f, whole, lane);
			continue;
		}

		// ---- GMSK main pass: lane owns outputs 5*lane .. 5*lane+4 with the untruncated composite ----
		const int nout = p.n_gmsk_soft;
		const int i0 = 5 * lane;
		float acc[5] = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
		if (i0 < nout) {
			// outputs 0,2,4 of the lane use one component, 1,3 the other (parity of i0 decides which)
			const float *uA = u + ((i0 & 1) ? kCompWords : 0) + i0;
			const float *uB = u + ((i0 & 1) ? 0 : kCompWords) + i0;
#pragma unroll
			for (int r = 0; r < 4; r++) {
				float xa[13], xb[11], c[9];
#pragma unroll
				for (int k = 0; k < 13; k++) xa[k] = uA[r * kPlane + k];
#pragma unroll
				for (int k = 0; k < 11; k++) xb[k] = uB[r * kPlane + 1 + k];
#pragma unroll
				for (int k = 0; k < 9; k++) c[k] = ce[4 * k + r];
#pragma unroll
				for (int k = 0; k < 9; k++) {
					acc[0] = fmaf(xa[k], c[k], acc[0]);
					acc[2] = fmaf(xa[k + 2], c[k], acc[2]);
					acc[4] = fmaf(xa[k + 4], c[k], acc[4]);
					acc[1] = fmaf(xb[k], c[k], acc[1]);
					acc[3] = fmaf(xb[k + 2], c[k], acc[3]);
				}
			}
		}
		float *orow = p.soft + (size_t)b * p.soft_stride;
		// leading outputs have their decimator taps truncated from below (recomputed next); outputs with
		// 4i > 624 + whole are truncated from above (generic path at the end)
		const int nlead = min(nout, (max(15, 15 + whole) + 3) >> 2);
		const int top = 624 + whole;
#pragma unroll
		for (int a = 0; a < 5; a++) {
			const int i = i0 + a;
			if (i < nout && i >= nlead && 4 * i <= top)
				orow[i] = derot_real(i, acc[a], acc[a]);
		}
		// ---- leading outputs: truncated composites, 4 lanes per output ----
		for (int base = 0; base < nlead; base += 8) {
			const int i = base + (lane >> 2), part = lane & 3;
			float a = 0.0f;
			int kmin = 16;
			if (i < nlead) {
				kmin = max(0, max(15 - 4 * i, 15 - 4 * i + whole));
				const int kmax = min(15, 639 + whole - 4 * i);
				if (kmin <= 15 && kmax == 15) {
					// taps 9*part .. 9*part+8 (36th tap of the row is zero padding)
					const float *__restrict__ c = p.comp + ((size_t)f * 16 + kmin) * 36 + 9 * part;
					const float *uc = u + ((i & 1) ? kCompWords : 0);
					const int w0 = 4 * i + e + 9 * part;
#pragma unroll
					for (int j = 0; j < 9; j++) {
						const int w = w0 + j;
						a = fmaf(uc[(w & 3) * kPlane + (w >> 2)], __ldg(&c[j]), a);
					}
				} else if (kmin <= kmax && part == 0) {
					const float2 d = slow_output(u, 4 * i + e, f, kmin, kmax);
					a = (i & 1) ? d.y : d.x;
				}
			}
			a += __shfl_xor_sync(0xffffffffu, a, 1);
			a += __shfl_xor_sync(0xffffffffu, a, 2);
			if (i < nlead && part == 0)
				orow[i] = derot_real(i, a, a);
		}
		// ---- outputs truncated from above (shifted burst runs past sample 624): rare, generic path ----
		if (4 * (nout - 1) > top) {
			for (int i = nlead + lane; i < nout; i += 32) {
				if (4 * i <= top) continue; // written by the main pass
				const int kmin = max(0, max(15 - 4 * i, 15 - 4 * i + whole));
				const int kmax = min(15, 639 + whole - 4 * i);
				float2 d = make_float2(0.0f, 0.0f);
				if (kmin <= kmax) d = slow_output(u, 4 * i + e, f, kmin, kmax);
				orow[i] = derot_real(i, d.x, d.y);
			}
		}
	}
}

} // namespace trxb200
