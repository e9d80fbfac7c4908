// demod.cu — batched soft demodulation (demodAnyBurst, sigProcLib.cpp:2130-2137) for sm_100a.
//
// Reference chain per burst (demodCommon :2030-2048 + demodGmskBurst :2055-2072):
//   delayVector(-toa*4): 20-tap fractional-delay FIR over all 625 samples + integer shift
//   scaleVector(1/amp), downsampleBurst: 16-tap decimating FIR (only 156 outputs are kept)
//   GMSKReverseRotate at 1 sps, real part.
// All of it is linear, and only every 4th sample of the delayed burst survives, so this kernel
// evaluates the 156 (or 148) surviving outputs directly with the 35-tap composite filter
// delay[f] (*) decimator built on the host (tables.cpp, `comp`), which is 3x fewer MACs than the
// two-stage form and reads each burst from HBM exactly once.  Where the reference's intermediate
// vectors are truncated (samples shifted in from outside the 625-sample burst are zero, history
// before sample 0 of the decimator is zero) the affected leading outputs use the composite
// truncated to the surviving decimator taps (comp[f][kmin]); outputs truncated from above take a
// generic two-stage path.  The complex gain 1/amp is applied while the burst is staged into shared
// memory, and because e^{-j*pi*n/2} only selects +-re/+-im (the reference's table has ~1e-16
// leakage from double-precision phase accumulation, far below the 1e-4 soft-bit tolerance) a GMSK
// output needs one real 35-tap dot product.  This chain feeds no decisions, so FMA is used.
//
// Shared memory layout per warp: the scaled, zero-padded burst window in polyphase-planar form
// u[comp][phase][q] (sample index = 4q+phase) so that lanes = consecutive outputs read consecutive
// words for every tap.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {
namespace {

constexpr int kPlane = 168;		   // floats per polyphase plane (>= 165)
constexpr int kCompWords = 4 * kPlane;	   // one component
constexpr int kWin = 4 * 156 + 36;	   // staged window length (660)
constexpr int kDemodWarpFloats = 2 * kCompWords + 2 * 164; // + complex dec[160+4] scratch for EDGE

__device__ __forceinline__ float win_get(const float *u, int comp, int bidx)
{
	return u[comp * kCompWords + (bidx & 3) * kPlane + (bidx >> 2)];
}

// generic two-stage evaluation of one decimated sample restricted to decimator taps [kmin,kmax]
// (used only when the shifted burst runs off the top of the 625-sample vector)
__device__ float2 slow_output(const float *u, int i, int f, int kmin, int kmax)
{
	float2 acc = make_float2(0.0f, 0.0f);
	for (int k = kmin; k <= kmax; k++) {
		float yr = 0.0f, yi = 0.0f;
		if (f < 64) {
			for (int j = 0; j < 20; j++) {
				const float h = c_tab.delay[f][j];
				yr = fmaf(win_get(u, 0, 4 * i + k + j), h, yr);
				yi = fmaf(win_get(u, 1, 4 * i + k + j), h, yi);
			}
		} else {
			yr = win_get(u, 0, 4 * i + k + 9);
			yi = win_get(u, 1, 4 * i + k + 9);
		}
		acc.x = fmaf(yr, c_tab.dnsamp[k], acc.x);
		acc.y = fmaf(yi, c_tab.dnsamp[k], acc.y);
	}
	return acc;
}

} // namespace

__global__ void __launch_bounds__(256)
demod_kernel(DemodParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	float *u = reinterpret_cast<float *>(smem_raw) + (size_t)warp * kDemodWarpFloats;
	float2 *decs = reinterpret_cast<float2 *>(u + 2 * kCompWords);

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
		const int off = -24 - whole; // window index 0 <-> burst sample `off`
		const bool edge = (rc == 5);

		// ---- stage s*x into the planar window, zero outside the burst; clip scan rides along ----
		float mx = 0.0f;
		__syncwarp();
		for (int w = lane; w < kWin; w += 32) {
			const int src = w + off;
			float2 v = make_float2(0.0f, 0.0f);
			if (src >= 0 && src < 625) {
				const float2 raw = __ldg(&x[src]);
				mx = fmaxf(mx, fmaxf(fabsf(raw.x), fabsf(raw.y)));
				v.x = fmaf(raw.x, s.x, -raw.y * s.y);
				v.y = fmaf(raw.x, s.y, raw.y * s.x);
			}
			u[(w & 3) * kPlane + (w >> 2)] = v.x;
			u[kCompWords + (w & 3) * kPlane + (w >> 2)] = v.y;
		}
		if (p.flags && p.fix_clip) {
			// samples the window did not cover (only for extreme shifts)
			for (int i = lane; i < 625; i += 32) {
				if (i - off < 0 || i - off >= kWin) {
					const float2 raw = __ldg(&x[i]);
					mx = fmaxf(mx, fmaxf(fabsf(raw.x), fabsf(raw.y)));
				}
			}
#pragma unroll
			for (int o = 16; o; o >>= 1)
				mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
			if (lane == 0 && mx > 30000.0f) p.flags[b] |= 4;
		}
		__syncwarp();

		const int nout = edge ? 156 : p.n_gmsk_soft;
		for (int i = lane; i < ((nout + 31) & ~31); i += 32) {
			float2 d = make_float2(0.0f, 0.0f);
			if (i < nout) {
				const int kmin = max(0, max(15 - 4 * i, 15 - 4 * i + whole));
				const int kmax = min(15, 639 + whole - 4 * i);
				if (kmin <= kmax) {
					if (kmax == 15) {
						const float *__restrict__ c = p.comp + ((size_t)f * 16 + kmin) * 36;
						if (edge) {
#pragma unroll
							for (int t = 0; t < 35; t++) {
								const float ct = __ldg(&c[t]);
								const int q = (t & 3) * kPlane + i + (t >> 2);
								d.x = fmaf(u[q], ct, d.x);
								d.y = fmaf(u[kCompWords + q], ct, d.y);
							}
						} else {
							// only the component e^{-j*pi*i/2} selects is needed
							const float *uc = u + ((i & 1) ? kCompWords : 0);
							float a = 0.0f;
#pragma unroll
							for (int t = 0; t < 35; t++)
								a = fmaf(uc[(t & 3) * kPlane + i + (t >> 2)], __ldg(&c[t]), a);
							if (i & 1) d.y = a; else d.x = a;
						}
					} else {
						d = slow_output(u, i, f, kmin, kmax);
					}
				}
			}
			if (!edge) {
				if (i < nout) {
					// GMSKReverseRotate(1 sps) + real part: Re(rrot1[i] * d)
					const float2 r = c_tab.rrot1[i];
					const float v = (i & 1) ? -(r.y * d.y) : (r.x * d.x);
					p.soft[(size_t)b * p.soft_stride + i] = v;
				}
			} else if (i < 156) {
				decs[2 + i] = d;
			}
		}

		if (edge) {
			// ---- demodEdgeBurst :2105-2128 on the 156 decimated samples ----
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
					rot = cmul_exact(make_float2(er, ei), c_tab.edge_derot[i & 15]); // derotateEdgeBurst :691-711
					if (i >= 8 && i < 148) {
						// computeEdgeCI :2074-2093
						const float step = 2.0f * 3.14159274f / 8.0f;
						int k = (int)roundf(atan2f(rot.y, rot.x) / step);
						k = min(max(k, -4), 4);
						const float2 ideal = c_tab.edge_ideal[k + 4];
						const float2 e = make_float2(fs(ideal.x, rot.x), fs(ideal.y, rot.y));
						err += norm2(e);
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
	}
}

} // namespace trxb200
