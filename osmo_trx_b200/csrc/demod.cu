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
// truncated from above take a generic two-stage path.  This chain feeds no decisions, so FMA is used.
//
// The FIR runs on the RAW complex samples with Blackwell's packed FP32 pipe: one FFMA2
// (fma.rn.f32x2) advances the (re,im) pair of an output by one real tap, so the complex sum costs the
// same issue slots a real one would, and the staging pass is a pure copy (no arithmetic).  The
// complex gain 1/amp and the e^{-j*pi*n/2} derotation (which only selects +-re/+-im; the table's
// ~1e-16 leakage is far below the 1e-4 soft-bit tolerance) are applied once per output.
//
// Work mapping, one warp per burst (the kernel is bound by the shared-memory data pipe, so every
// window sample is written to and read from shared memory exactly once):
//   stage   16-byte global loads (two samples) -> 16-byte shared stores into a 680-sample window; 16-byte
//           slot s lives at s ^ ((s/40)&1), which keeps both the staging stores and the per-lane reads
//           below (lane stride 10 slots) bank-conflict free; the window origin is aligned to the 16-byte
//           grid of the row, the residual shift e is folded into the tap index.  The next burst of the
//           warp is prefetched into L2 meanwhile.
//   FIR     transposed form: lane l owns window samples 20l..20l+19 (10 LDS.128) and scatters each into
//           the 13 outputs 5l-8..5l+4 it can reach (180 FFMA2 with the tap as scalar-broadcast operand,
//           taps fetched warp-uniformly from __constant__ comp0[f][e]); the 8 partial sums that belong to
//           lanes l-1 and l-2 are handed over with shuffles once per burst.
//   edges   the few leading outputs whose decimator taps are truncated get the dropped terms subtracted:
//           the (at most 31) delayed samples Y[v] those taps would have read are evaluated one per lane
//           (20-tap fractional filter from __constant__), then 4 lanes per output sum g[k]*Y[4i+k], k < kmin.
//   store   outputs are staged in shared memory and written with 16-byte coalesced stores.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {
namespace {

constexpr int kPairs = 340;		 // 16-byte sample pairs (slots) in the window: samples 0 .. 679
constexpr int kScratchFloats = 2 * 164;	 // output staging + Y scratch (GMSK) / complex decimated samples (EDGE)
constexpr int kDemodWarpFloats = 4 * kPairs + kScratchFloats;
constexpr int kYOff = 192;		 // float offset of the Y scratch (float2[32]) inside the scratch area

// ---- packed FP32 (sm_100 FFMA2): both halves are IEEE fma.rn ----
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
	unsigned long long ra, rb, rc, rd;
	ra = *reinterpret_cast<unsigned long long *>(&a);
	rb = *reinterpret_cast<unsigned long long *>(&b);
	rc = *reinterpret_cast<unsigned long long *>(&c);
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
	return *reinterpret_cast<float2 *>(&rd);
}

__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
	unsigned long long ra, rb, rd;
	ra = *reinterpret_cast<unsigned long long *>(&a);
	rb = *reinterpret_cast<unsigned long long *>(&b);
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
	return *reinterpret_cast<float2 *>(&rd);
}

// physical 16-byte slot of logical slot s (window samples 2s, 2s+1)
__device__ __forceinline__ int slot_phys(int s) { return s ^ ((s / 40) & 1); }

__device__ __forceinline__ float2 win_get(const float2 *U, int w)
{
	return U[2 * slot_phys(w >> 1) + (w & 1)];
}

// generic two-stage evaluation of one decimated sample restricted to decimator taps [kmin,kmax]
// (only when the shifted burst runs off the top of the 625-sample vector); w0 = window index of tap 0
__device__ __noinline__ float2 slow_output(const float2 *U, int w0, int f, int kmin, int kmax)
{
	float2 acc = make_float2(0.0f, 0.0f);
	for (int k = kmin; k <= kmax; k++) {
		float2 y = make_float2(0.0f, 0.0f);
		if (f < 64) {
			for (int j = 0; j < 20; j++) {
				const float h = c_tab.delay[f][j];
				y = ffma2(win_get(U, w0 + k + j), make_float2(h, h), y);
			}
		} else {
			y = win_get(U, w0 + k + 9);
		}
		const float g = c_tab.dnsamp[k];
		acc = ffma2(y, make_float2(g, g), acc);
	}
	return acc;
}

// (complex)1/amp applied to an unscaled FIR sum
__device__ __forceinline__ float2 cscale(float2 a, float2 s)
{
	return make_float2(fmaf(a.x, s.x, -a.y * s.y), fmaf(a.x, s.y, a.y * s.x));
}

// Re(e^{-j*pi*i/2} * s * a): scale + GMSKReverseRotate(1 sps) + real part (sigProcLib.cpp:262-287,2011-2022)
__device__ __forceinline__ float soft_out(int i, float2 a, float2 s)
{
	const float v = (i & 1) ? fmaf(a.x, s.y, a.y * s.x) : fmaf(a.x, s.x, -a.y * s.y);
	return (i & 2) ? -v : v;
}

// one output with a (possibly truncated) composite: kmin..15 decimator taps present, 35 composite taps
__device__ __forceinline__ float2 comp_output(const DemodParams &p, const float2 *U, int i, int e, int f, int kmin)
{
	const float *__restrict__ c = p.comp + ((size_t)f * 16 + kmin) * 36;
	float2 d = make_float2(0.0f, 0.0f);
#pragma unroll 5
	for (int t = 0; t < 35; t++) {
		const float ct = __ldg(&c[t]);
		d = ffma2(win_get(U, 4 * i + t + e), make_float2(ct, ct), d);
	}
	return d;
}

// ---- EDGE: demodEdgeBurst :2105-2128 on the staged window (complex outputs) ----
__device__ __noinline__ void demod_edge_burst(const DemodParams &p, int b, const float2 *U, float2 *decs, float2 s, int e, int f,
					      int whole, int lane)
{
	for (int i = lane; i < 160; i += 32) {
		float2 d = make_float2(0.0f, 0.0f);
		if (i < 156) {
			const int kmin = max(0, max(15 - 4 * i, 15 - 4 * i + whole));
			const int kmax = min(15, 639 + whole - 4 * i);
			if (kmin <= kmax) {
				if (kmax == 15)
					d = comp_output(p, U, i, e, f, kmin);
				else
					d = slow_output(U, 4 * i + e, f, kmin, kmax);
			}
			decs[2 + i] = cscale(d, s);
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

__device__ __forceinline__ void prefetch_row_l2(const float2 *x, int lane)
{
	// 625 samples = 5000 B: one bulk L2 prefetch of the 16-byte aligned span inside the row
	if (lane == 0) {
		const uintptr_t a = (reinterpret_cast<uintptr_t>(x) + 15u) & ~(uintptr_t)15u;
		asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(4976));
	}
}

} // namespace

__global__ void __launch_bounds__(256, 3)
demod_kernel(DemodParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	float2 *U = reinterpret_cast<float2 *>(smem_raw) + (size_t)warp * (kDemodWarpFloats / 2);
	float *ostage = reinterpret_cast<float *>(U + 2 * kPairs);
	float2 *decs = reinterpret_cast<float2 *>(ostage);
	float2 *yv = reinterpret_cast<float2 *>(ostage + kYOff);
	const unsigned base_par = (unsigned)((reinterpret_cast<uintptr_t>(p.bursts) >> 3) & 1u);
	const int step = gridDim.x * wpb;
	// decimator taps this lane applies in the leading-output correction (k = 4*(lane&3) + kk)
	float gk[4];
#pragma unroll
	for (int kk = 0; kk < 4; kk++) gk[kk] = p.dnsamp_g[4 * (lane & 3) + kk];

	int b = blockIdx.x * wpb + warp;
	// per-burst scalars are fetched one burst ahead
	int rc_n = 0;
	float2 amp_n = make_float2(1.0f, 0.0f);
	float toa_n = 0.0f;
	if (b < p.n) {
		rc_n = p.rc[b];
		amp_n = reinterpret_cast<const float2 *>(p.amp)[b];
		toa_n = p.toa[b];
	}
	for (; b < p.n; b += step) {
		const int rc = rc_n;
		const float2 amp = amp_n;
		const float toa = toa_n;
		const float2 *x = reinterpret_cast<const float2 *>(p.bursts) + (size_t)b * p.stride;
		const int bn = b + step;
		if (bn < p.n) {
			rc_n = p.rc[bn];
			amp_n = reinterpret_cast<const float2 *>(p.amp)[bn];
			toa_n = p.toa[bn];
			if (rc_n > 0)
				prefetch_row_l2(reinterpret_cast<const float2 *>(p.bursts) + (size_t)bn * p.stride, lane);
		}

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
		const float an = norm2(amp);
		const float2 s = make_float2(amp.x / an, -amp.y / an); // (complex)1.0 / amp
		const float delay = fm(-toa, 4.0f);
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
		// ---- stage the raw burst into the window, zero outside the burst ----
		// All 16-byte loads of the burst are issued before the first use (11 in flight per lane) so the warp
		// pays the memory latency once per burst.  Slot s starts at burst sample p0 = off2 + 2s, 16-byte
		// aligned by construction; the one slot that straddles an end of the burst is patched below.
		{
			float4 ld[11];
#pragma unroll
			for (int it = 0; it < 11; it++) {
				const int sl = lane + 32 * it;
				const int p0 = off2 + 2 * sl;
				ld[it] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				if ((it < 10 || sl < kPairs) && (unsigned)p0 <= 623u)
					ld[it] = __ldg(reinterpret_cast<const float4 *>(x + p0));
			}
			float4 *U4 = reinterpret_cast<float4 *>(U);
#pragma unroll
			for (int it = 0; it < 11; it++) {
				const int sl = lane + 32 * it;
				if (it < 10 || sl < kPairs)
					U4[slot_phys(sl)] = ld[it];
			}
			// the straddling slot: samples (-1, 0) or (624, 625), one per burst depending on the row parity
			if (lane == 0) {
				if (((-1 - off2) & 1) == 0) {
					const int sl = (-1 - off2) >> 1;
					if (sl >= 0 && sl < kPairs) U[2 * slot_phys(sl) + 1] = __ldg(&x[0]);
				} else {
					const int sl = (624 - off2) >> 1;
					if (sl >= 0 && sl < kPairs) U[2 * slot_phys(sl)] = __ldg(&x[624]);
				}
			}
		}
		__syncwarp();

		if (edge) {
			demod_edge_burst(p, b, U, decs, s, e, f, whole, lane);
			continue;
		}

		// ---- GMSK main pass (transposed FIR): lane owns window samples 20*lane .. 20*lane+19 and
		//      accumulates into outputs i = 5*lane - 8 + m, m = 0..12; sample j meets output m with
		//      tap u = j + 32 - 4m (0 <= u <= 35), coefficient ce[u] = comp0[f][e][u] ----
		const int nout = p.n_gmsk_soft;
		{
			float2 acc[13];
#pragma unroll
			for (int m = 0; m < 13; m++) acc[m] = make_float2(0.0f, 0.0f);
			const float4 *U4 = reinterpret_cast<const float4 *>(U);
			const int sw = (lane >> 2) & 1;
			const float *__restrict__ ce = c_tab.comp0[f][e];
#pragma unroll
			for (int h = 0; h < 2; h++) {
				// samples j = 4k + 2h and 4k + 2h + 1, k = 0..4
				float2 xa[5], xb[5];
#pragma unroll
				for (int k = 0; k < 5; k++) {
					const float4 v = U4[(10 * lane + 2 * k + h) ^ sw];
					xa[k] = make_float2(v.x, v.y);
					xb[k] = make_float2(v.z, v.w);
				}
#pragma unroll
				for (int g = 0; g < 9; g++) {
					// taps u = 4g + 2h (for xa) and u + 1 (for xb)
					const float ca = ce[4 * g + 2 * h], cb = ce[4 * g + 2 * h + 1];
#pragma unroll
					for (int k = 0; k < 5; k++) {
						// j = 4k + 2h, u = 4g + 2h  =>  m = (j + 32 - u) / 4 = k + 8 - g
						const int m = k + 8 - g;
						acc[m] = ffma2(xa[k], make_float2(ca, ca), acc[m]);
						acc[m] = ffma2(xb[k], make_float2(cb, cb), acc[m]);
					}
				}
			}
			// hand the partial sums of outputs owned by lanes l-1 (m = 3..7) and l-2 (m = 0..2) over
			float2 fin[5];
#pragma unroll
			for (int a = 0; a < 5; a++) {
				float2 t1;
				t1.x = __shfl_down_sync(0xffffffffu, acc[3 + a].x, 1);
				t1.y = __shfl_down_sync(0xffffffffu, acc[3 + a].y, 1);
				fin[a] = fadd2(acc[8 + a], t1);
				if (a >= 2) {
					float2 t2;
					t2.x = __shfl_down_sync(0xffffffffu, acc[a - 2].x, 2);
					t2.y = __shfl_down_sync(0xffffffffu, acc[a - 2].y, 2);
					fin[a] = fadd2(fin[a], t2);
				}
			}
			// soft value = Re(z_i * sum), z_i = (1/amp) * (-j)^i, i = 5*lane + a  (i mod 4 = (lane + a) mod 4);
			// lanes 30, 31 lack their right-hand neighbours: outputs >= 150 are finished by the generic path
			if (lane < 30) {
				float zx = (lane & 1) ? s.y : s.x, zy = (lane & 1) ? -s.x : s.y;
				if (lane & 2) { zx = -zx; zy = -zy; }
#pragma unroll
				for (int a = 0; a < 5; a++) {
					ostage[5 * lane + a] = fmaf(zx, fin[a].x, -zy * fin[a].y);
					const float t = zx; // z *= -j
					zx = zy;
					zy = -t;
				}
			}
		}
		// leading outputs have their decimator taps truncated from below (corrected next); outputs with
		// 4i > 624 + whole are truncated from above, outputs >= 150 lack window lanes (generic path below)
		const int nlead = min(nout, (max(15, 15 + whole) + 3) >> 2);
		const int top = 624 + whole;
		const int nv = 15 + max(0, whole); // delayed samples Y[v], v < nv, are what the dropped taps would read
		__syncwarp();
		if (nv <= 32) {
			// Y[v] = sum_j win(v + e + j) * delay[f][j]: the delayed sample decimator tap k of output i reads, v = 4i + k
			float2 y = make_float2(0.0f, 0.0f);
			if (f < 64) {
#pragma unroll
				for (int j = 0; j < 20; j++) {
					const float hj = c_tab.delay[f][j];
					y = ffma2(U[lane + e + j], make_float2(hj, hj), y); // w < 80: slot_phys is the identity
				}
			} else {
				y = U[lane + e + 9];
			}
			yv[lane] = y;
			__syncwarp();
			for (int base = 0; base < nlead; base += 8) {
				const int i = base + (lane >> 2), part = lane & 3;
				const int kmin = min(16, max(0, max(15 - 4 * i, 15 - 4 * i + whole)));
				float2 d = make_float2(0.0f, 0.0f);
				if (i < nlead) {
#pragma unroll
					for (int kk = 0; kk < 4; kk++) {
						const int k = 4 * part + kk;
						if (k < kmin)
							d = ffma2(yv[4 * i + k], make_float2(gk[kk], gk[kk]), d);
					}
				}
				d.x += __shfl_xor_sync(0xffffffffu, d.x, 1);
				d.y += __shfl_xor_sync(0xffffffffu, d.y, 1);
				d.x += __shfl_xor_sync(0xffffffffu, d.x, 2);
				d.y += __shfl_xor_sync(0xffffffffu, d.y, 2);
				if (i < nlead && part == 0)
					ostage[i] -= soft_out(i, d, s);
			}
		}
		// ---- generic per-output path (rare): outputs truncated from above (shifted burst runs past sample
		//      624), outputs beyond the main pass (>= 150), leading outputs of very early bursts (nv > 32) ----
		if (4 * (nout - 1) > top || nout > 150 || nv > 32) {
			for (int i = lane; i < nout; i += 32) {
				if (4 * i <= top && i < 150 && !(nv > 32 && i < nlead)) continue; // main pass result stands
				const int kmin = max(0, max(15 - 4 * i, 15 - 4 * i + whole));
				const int kmax = min(15, 639 + whole - 4 * i);
				float2 d = make_float2(0.0f, 0.0f);
				if (kmin <= kmax)
					d = (kmax == 15) ? comp_output(p, U, i, e, f, kmin) : slow_output(U, 4 * i + e, f, kmin, kmax);
				ostage[i] = soft_out(i, d, s);
			}
		}
		__syncwarp();
		// ---- coalesced store of the soft row ----
		float *orow = p.soft + (size_t)b * p.soft_stride;
		if (((reinterpret_cast<uintptr_t>(orow) & 15u) == 0) && (nout & 3) == 0) {
			const float4 *os4 = reinterpret_cast<const float4 *>(ostage);
			for (int j = lane; j < (nout >> 2); j += 32)
				reinterpret_cast<float4 *>(orow)[j] = os4[j];
		} else {
			for (int j = lane; j < nout; j += 32)
				orow[j] = ostage[j];
		}
	}
}

} // namespace trxb200
