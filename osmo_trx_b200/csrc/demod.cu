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
// truncated from above get the dropped terms subtracted.  This chain feeds no decisions, so FMA is used.
//
// The FIR runs on the RAW complex samples with Blackwell's packed FP32 pipe: one FFMA2
// (fma.rn.f32x2) advances the (re,im) pair of an output by one real tap, so the complex sum costs the
// same issue slots a real one would, and the staging pass is a pure copy (no arithmetic).  The
// complex gain 1/amp and the e^{-j*pi*n/2} derotation (which only selects +-re/+-im; the table's
// ~1e-16 leakage is far below the 1e-4 soft-bit tolerance) are applied once per output.
//
// Work mapping, one warp per burst (the kernel is bound by the shared-memory data pipe, so every
// window sample is written to and read from shared memory exactly once):
//   stage   TMA bulk copies (cp.async.bulk global -> shared, completion on a per-warp mbarrier) bring the
//           burst into a 680-sample window, double buffered: the copies of the warp's NEXT burst are in
//           flight while the current one is filtered, so no warp ever waits on HBM latency and the
//           staging costs neither registers nor load/store-pipe wavefronts.  One bulk copy per burst
//           (issued by lane 0) into a linear window of 340 16-byte slots; the per-lane 16-byte reads
//           below (lane stride 10 slots) are 2-way bank conflicted, which the shared-memory pipe has room
//           for, whereas a padded layout would need nine copies and a uniform-register loop to issue them.
//           The window origin is aligned to the 16-byte grid of the row, the residual shift e is folded
//           into the tap index; slots outside the burst are zero-filled, the one sample pair straddling
//           an end of the burst is patched with an ordinary 8-byte load.
//   FIR     transposed form: lane l owns window samples 20l..20l+19 (10 LDS.128) and scatters each into
//           the 13 outputs 5l-8..5l+4 it can reach (180 FFMA2 with the tap as scalar-broadcast operand,
//           taps fetched warp-uniformly from __constant__ comp0[f][e]); the 8 partial sums that belong to
//           lanes l-1 and l-2 are handed over with shuffles once per burst.
//   edges   the few leading outputs whose decimator taps are truncated are evaluated again, directly, with the composite
//           truncated to the surviving taps (comp[f][kmin], global memory / L2): 4 lanes per output, 9 taps each.  Trailing
//           outputs of very late bursts get the dropped terms subtracted instead: the (at most 16) delayed samples those
//           taps would have read are evaluated one per lane (20-tap fractional filter from __constant__).
//   store   outputs are staged in shared memory and written with 16-byte coalesced stores.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {
namespace {

constexpr int kPairs = 340;		 // 16-byte sample pairs (slots) in the window: samples 0 .. 679
constexpr int kBufSlots = kPairs;	 // one window buffer (linear)
constexpr int kYOff = 320;		 // float offset of the leading-edge Y scratch (float2[32]); outputs (float[156] GMSK /
					 // float2[160] EDGE) sit below it
constexpr int kYTopOff = kYOff + 64;	 // float offset of the trailing-edge Y scratch (float2[16])
constexpr int kYxOff = kYTopOff + 64;	 // float offset of the converted head of an int16 window (float2[56]); >= 444: the EDGE soft
					 // row is staged over the area below
constexpr int kIdealOff = kYxOff + 112;	 // float offset of the warp's copy of the 9 ideal 8-PSK points (float2[9], padded to 32 floats)
constexpr int kScratchFloats = kIdealOff + 32;
constexpr int kDemodWarpFloats = 2 * 4 * kBufSlots + kScratchFloats + 4; // 2 window buffers + scratch + 2 mbarriers

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
__device__ __forceinline__ int slot_phys(int s) { return s; }

// ---- TMA bulk-copy primitives (smem_u32 / mbar_init / mbar_wait / mbar_arrive are detect.cu's) ----
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
		     "r"(bytes), "r"(bar)
		     : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// window sample w as complex float.  I16: the window holds the radio's int16 I/Q pairs as they arrived (pull path:
// convert_short_float, arch/x86/convert.c:37-79, happens here, on the way to the FIR; (float)int16 is exact)
template <bool I16>
__device__ __forceinline__ float2 win_get(const float2 *U, int w)
{
	if constexpr (I16) return cvt_s2(reinterpret_cast<const unsigned *>(U)[w]);
	else return U[w];
}

// generic two-stage evaluation of one decimated sample restricted to decimator taps [kmin,kmax]
// (only when the shifted burst runs off the top of the 625-sample vector); w0 = window index of tap 0
template <bool I16>
__device__ __noinline__ float2 slow_output(const float2 *U, int w0, int f, int kmin, int kmax)
{
	float2 acc = make_float2(0.0f, 0.0f);
	for (int k = kmin; k <= kmax; k++) {
		float2 y = make_float2(0.0f, 0.0f);
		if (f < 64) {
			for (int j = 0; j < 20; j++) {
				const float h = c_tab.delay[f][j];
				y = ffma2(win_get<I16>(U, w0 + k + j), make_float2(h, h), y);
			}
		} else {
			y = win_get<I16>(U, w0 + k + 9);
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

// complex product with fused multiply-adds (non-decision chains)
__device__ __forceinline__ float2 cmul_fast(float2 a, float2 b)
{
	return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// Re(e^{-j*pi*i/2} * s * a): scale + GMSKReverseRotate(1 sps) + real part (sigProcLib.cpp:262-287,2011-2022)
__device__ __forceinline__ float soft_out(int i, float2 a, float2 s)
{
	const float v = (i & 1) ? fmaf(a.x, s.y, a.y * s.x) : fmaf(a.x, s.x, -a.y * s.y);
	return (i & 2) ? -v : v;
}

// soft values staged in shared memory -> the datagram's soft bytes (vectorSlicer + trxd_fill_burst_normalized255), written
// to the slot's row behind the header.  bytes: scratch indexed by (packet byte - 8); rows that cannot hold the burst
// are left untouched (header_kernel flags them).
__device__ __forceinline__ void store_soft_bytes(const DemodParams &p, int b, const float *soft, int nbits, uint8_t *bytes, int lane)
{
	const int hdr = p.pkt_hdr;
	const int len = hdr + nbits + (p.pkt_v0 ? 2 : 0);
	if (len > p.pkt_stride) return;
	for (int j = lane; j < nbits; j += 32) bytes[hdr - 8 + j] = (uint8_t)soft_to_u8(soft[j]);
	if (p.pkt_v0 && lane < 2) bytes[hdr - 8 + nbits + lane] = 0;
	if (lane < hdr - 8) bytes[lane] = 0; // v1: bytes 8..10 are header bytes (header_kernel fills them afterwards)
	__syncwarp();
	uint8_t *row = p.pkt + (size_t)b * p.pkt_stride;
	if ((reinterpret_cast<uintptr_t>(row) & 3u) == 0) {
		// whole words from byte 8 (v1: its bytes 8..10 are rewritten by header_kernel, which runs after this kernel)
		const uint32_t *s4 = reinterpret_cast<const uint32_t *>(bytes);
		const int nw = (len - 8) >> 2;
		for (int k = lane; k < nw; k += 32) reinterpret_cast<uint32_t *>(row + 8)[k] = s4[k];
		if (lane < ((len - 8) & 3)) row[8 + 4 * nw + lane] = bytes[4 * nw + lane];
	} else {
		for (int k = hdr + lane; k < len; k += 32) row[k] = bytes[k - 8];
	}
}

// ---- EDGE: demodEdgeBurst :2105-2128 after the decimator: decs[2 + i], i < 156, hold the scaled complex
//      1-sps samples (the shared FIR pass below produced them) ----
// rot_lane: the lane's derotation factor edge_tab[lane & 15] (symbol i = lane + 32 r has i & 15 = lane & 15); ideal: the warp's
// shared-memory copy of the nine ideal points - neither is fetched from global memory per burst
// size: length of the 1-sps vector (156 after the 4:1 decimator; the burst's own length, 156 or 157, on the 1-sps path):
// computeEdgeCI walks symbols 8 .. size - 9 and divides by size - 16
__device__ __noinline__ void demod_edge_tail(const DemodParams &p, int b, float2 *decs, int lane, float2 rot_lane, const float2 *ideal_s, int size = 156)
{
	if (lane < 2) { decs[lane] = make_float2(0.0f, 0.0f); decs[2 + size + lane] = make_float2(0.0f, 0.0f); }
	__syncwarp();
	float err = 0.0f;
	float2 rot[5];
	float c0[5];
#pragma unroll
	for (int k = 0; k < 5; k++) c0[k] = c_tab.c0_inv[k];
	// none of this chain feeds a decision (soft bits and C/I carry the 1e-4 tolerance): FMA forms throughout
#pragma unroll
	for (int r = 0; r < 5; r++) {
		const int i = lane + 32 * r;
		rot[r] = make_float2(0.0f, 0.0f);
		if (i < max(148, size - 8)) { // softSliceEdgeBurst consumes symbols 0..147 only, computeEdgeCI 8 .. size - 9
			// 5-tap static equaliser, NO_DELAY span (convolve_base.c:27-60)
			float2 eq = make_float2(0.0f, 0.0f);
#pragma unroll
			for (int k = 0; k < 5; k++) eq = ffma2(decs[i + k], make_float2(c0[k], c0[k]), eq);
			rot[r] = cmul_fast(eq, rot_lane); // derotateEdgeBurst :691-711
			if (i >= 8 && i < size - 8) {
				// computeEdgeCI :2074-2093: squared distance to the nearest ideal 8-PSK point.  The reference picks the point as
				// round(atan2(y, x) / (pi/4)); on a circle the point nearest in angle is the point nearest in distance, and
				// the constellation is symmetric under |x|, |y| and their exchange: with hi = max(|x|, |y|), lo = min(|x|, |y|)
				// the candidates are (1, 0) and (c, c), c = cos(pi/4).  (The table's points carry cosf's 4e-8 leakage, far
				// below the 1e-4 tolerance of this chain.)
				const float ax = fabsf(rot[r].x), ay = fabsf(rot[r].y);
				const float hi = fmaxf(ax, ay), lo = fminf(ax, ay);
				const float a1 = hi - 1.0f, b1 = hi - 0.70710678f, b2 = lo - 0.70710678f;
				err += fminf(fmaf(a1, a1, lo * lo), fmaf(b1, b1, b2 * b2));
			}
		}
	}
	__syncwarp(); // every lane has read its decimated samples: the area becomes the soft-row staging buffer
	// softSliceEdgeBurst :1962-2006
	float *ost = reinterpret_cast<float *>(decs);
	const float2 rot1 = c_tab.edge_rot1, rot2 = c_tab.edge_rot2;
#pragma unroll
	for (int r = 0; r < 5; r++) {
		const int i = lane + 32 * r;
		if (i < 148) {
			const float2 r1 = cmul_fast(rot[r], rot1);
			const float r2y = fmaf(fabsf(r1.x), rot2.y, fabsf(r1.y) * rot2.x);
			ost[3 * i] = -r1.y;
			ost[3 * i + 1] = r1.x;
			ost[3 * i + 2] = -r2y;
		}
	}
	__syncwarp();
	if (p.pkt) {
		// pull chain: datagram bytes instead of the float row (scratch behind the 444 staged values)
		store_soft_bytes(p, b, ost, 444, reinterpret_cast<uint8_t *>(ost + kYxOff), lane);
		__syncwarp();
	} else {
	const int nvals = 3 * min(148, p.soft_stride / 3); // a row shorter than 444 values is never overrun
	float *orow = p.soft + (size_t)b * p.soft_stride;
	if ((reinterpret_cast<uintptr_t>(orow) & 15u) == 0 && (nvals & 3) == 0) {
		const float4 *os4 = reinterpret_cast<const float4 *>(ost);
		float4 *o4 = reinterpret_cast<float4 *>(orow);
		const int n4 = nvals >> 2; // <= 111
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int j = lane + 32 * k;
			if (j < n4) o4[j] = os4[j];
		}
	} else {
		for (int j = lane; j < nvals; j += 32)
			orow[j] = ost[j];
	}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1)
		err += __shfl_xor_sync(0xffffffffu, err, o);
	if (lane == 0)
		p.ci[b] = 3.0103f * __log2f(__fdividef((float)(size - 16), err));
}

} // namespace

// window geometry of a detected burst (demodCommon / delayVector :1046-1060)
struct BurstGeom {
	int whole, f, e, off2;
};
// phase: position of the row's first sample on the 16-byte grid, in samples (float rows: 0..1, int16 rows: 0..3);
// the window origin off2 is moved down to the grid, the residual shift e is folded into the tap index
template <bool I16>
__device__ __forceinline__ BurstGeom burst_geom(float toa, unsigned phase)
{
	BurstGeom g;
	const float delay = fm(-toa, 4.0f);
	g.whole = (int)floorf(delay);
	const float frac = fs(delay, (float)g.whole);
	g.f = 64;
	if ((double)fabsf(frac) > 1e-2) {
		g.f = (int)floorf(fm(frac, 64.0f));
		g.f = min(max(g.f, 0), 63);
	}
	const int off = -24 - g.whole;			  // window sample 0 <-> burst sample `off` (before alignment)
	g.e = (int)(((unsigned)off + phase) & (I16 ? 3u : 1u)); // residual shift so that off2 sits on the row's 16-byte grid
	g.off2 = off - g.e;				  // window sample w <-> burst sample w + off2; output i tap t reads w = 4i+t+e
	return g;
}

// Stage burst row x into window buffer U (async): zero-fill the slots outside the burst, launch the bulk copy.  Called by the whole warp; bar is the buffer's mbarrier.  The one sample pair that straddles
// an end of the burst cannot be bulk-copied (8-byte aligned only): lane 31 fetches it here and returns it in
// (patch, patch_idx); the caller stores it right before the buffer is consumed, so its latency is hidden too.
__device__ __forceinline__ void stage_async(const float2 *x, int off2, float2 *U, unsigned bar, int lane, float2 &patch,
					    int &patch_idx)
{
	// slot s holds burst samples off2 + 2s, off2 + 2s + 1; fully inside the burst for s_first <= s <= s_last
	const int s_first = off2 >= 0 ? 0 : ((-off2 + 1) >> 1);
	const int s_last = min(kPairs - 1, (623 - off2) >> 1); // floor; negative when the window lies beyond the burst
	const int count = max(0, s_last - s_first + 1);
	patch_idx = -1;
	if (lane == 31) {
		// the straddling slot: samples (-1, 0) or (624, 625), one per burst depending on the row parity
		if (((-1 - off2) & 1) == 0) {
			const int sl = (-1 - off2) >> 1;
			if (sl >= 0 && sl < kPairs) { patch = __ldg(&x[0]); patch_idx = 2 * slot_phys(sl) + 1; }
		} else {
			const int sl = (624 - off2) >> 1;
			if (sl >= 0 && sl < kPairs) { patch = __ldg(&x[624]); patch_idx = 2 * slot_phys(sl); }
		}
	}
	fence_proxy_async(); // this buffer's earlier generic-proxy accesses are ordered before the async writes below
	__syncwarp();
	float4 *U4 = reinterpret_cast<float4 *>(U);
	const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	const int nz_hi = count > 0 ? kPairs - 1 - s_last : kPairs; // zero slots above the burst (all of them if no overlap)
	const int nz_lo = count > 0 ? s_first : 0;
	if (nz_lo + nz_hi <= 32) {
		// the usual case: one predicated store per lane covers both ends
		const int sidx = lane < nz_lo ? lane : kPairs - nz_hi + (lane - nz_lo);
		if (lane < nz_lo + nz_hi) U4[slot_phys(sidx)] = z;
	} else {
		for (int sidx = lane; sidx < nz_lo; sidx += 32) U4[slot_phys(sidx)] = z;
		for (int sidx = kPairs - nz_hi + lane; sidx < kPairs; sidx += 32) U4[slot_phys(sidx)] = z;
	}
	__syncwarp();
	if (lane == 0) {
		mbar_arrive_expect_tx(bar, (unsigned)count * 16u);
		if (count > 0) bulk_g2s(smem_u32(U4 + s_first), x + off2 + 2 * s_first, (unsigned)count * 16u, bar);
	}
}

// int16 rows: a 16-byte slot holds FOUR samples, so up to three samples straddle either end of the burst; lanes
// 24..27 fetch the head slot's samples, lanes 28..31 the tail slot's (patch = the raw 32-bit I/Q word in .x).
constexpr int kSlots16 = 171; // 684 samples
__device__ __forceinline__ void stage_async16(const short2 *x, int off2, float2 *U, unsigned bar, int lane, float2 &patch,
					      int &patch_idx)
{
	// slot s holds burst samples off2 + 4s .. off2 + 4s + 3; fully inside the burst for s_first <= s <= s_last
	const int s_first = off2 >= 0 ? 0 : ((-off2 + 3) >> 2);
	const int s_last = min(kSlots16 - 1, (621 - off2) >> 2); // floor; negative when the window lies beyond the burst
	const int count = max(0, s_last - s_first + 1);
	patch_idx = -1;
	if (lane >= 24) {
		const int sl = (lane < 28) ? s_first - 1 : s_last + 1;
		const int w = 4 * sl + (lane & 3), nidx = off2 + w;
		if (sl >= 0 && sl < kSlots16 && nidx >= 0 && nidx <= 624) {
			patch.x = __int_as_float(__ldg(reinterpret_cast<const int *>(x) + nidx));
			patch_idx = w;
		}
	}
	fence_proxy_async();
	__syncwarp();
	float4 *U4 = reinterpret_cast<float4 *>(U);
	const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	const int nz_hi = count > 0 ? kSlots16 - 1 - s_last : kSlots16;
	const int nz_lo = count > 0 ? s_first : 0;
	if (nz_lo + nz_hi <= 32) {
		const int sidx = lane < nz_lo ? lane : kSlots16 - nz_hi + (lane - nz_lo);
		if (lane < nz_lo + nz_hi) U4[sidx] = z;
	} else {
		for (int sidx = lane; sidx < nz_lo; sidx += 32) U4[sidx] = z;
		for (int sidx = kSlots16 - nz_hi + lane; sidx < kSlots16; sidx += 32) U4[sidx] = z;
	}
	__syncwarp();
	if (lane == 0) {
		mbar_arrive_expect_tx(bar, (unsigned)count * 16u);
		if (count > 0) bulk_g2s(smem_u32(U4 + s_first), x + off2 + 4 * s_first, (unsigned)count * 16u, bar);
	}
}

// The terms of energyDetect(burst, 20 * sps) (pullRadioVector, Transceiver.cpp:723-731, sigProcLib.cpp:1573-1585) on
// an int16 slot: |x[4i]|^2, i < 80, written to the slot's row of p.pw.  The sequential float sum over them is
// header_kernel's (lanes = slots there: 80 dependent adds cost 2.5 instructions per slot instead of 100 here).
// The samples come from the staged window when it covers them (Uw != nullptr), else from the row.
__device__ __forceinline__ void energy_terms16(const short2 *xg, const short2 *Uw, int off2, int lane, float *dst)
{
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const int idx = lane + 32 * k;
		if (idx < 80) {
			const unsigned v = Uw ? reinterpret_cast<const unsigned *>(Uw)[4 * idx - off2]
					      : __ldg(reinterpret_cast<const unsigned *>(xg) + 4 * idx);
			dst[idx] = norm2(cvt_s2(v));
		}
	}
}

// the warp's scratch behind its two window buffers, and the decimator taps this lane applies in the edge corrections
struct DemodWarp {
	float *ostage;
	float2 *decs, *yv, *ytop, *yx;
	float gk[4]; // k = 4*(lane&3) + kk
	float2 rot_lane;     // edge_tab[lane & 15]
	const float2 *ideal; // shared-memory copy of edge_tab[16 .. 24]
};
__device__ __forceinline__ DemodWarp demod_warp_setup(const DemodParams &p, float *ostage, int lane)
{
	DemodWarp W;
	W.ostage = ostage;
	W.decs = reinterpret_cast<float2 *>(ostage);
	W.yv = reinterpret_cast<float2 *>(ostage + kYOff);
	W.ytop = reinterpret_cast<float2 *>(ostage + kYTopOff);
	W.yx = reinterpret_cast<float2 *>(ostage + kYxOff);
#pragma unroll
	for (int kk = 0; kk < 4; kk++) W.gk[kk] = p.dnsamp_g[4 * (lane & 3) + kk];
	W.rot_lane = p.edge_tab[lane & 15];
	float2 *ideal = reinterpret_cast<float2 *>(ostage + kIdealOff);
	if (lane < 9) ideal[lane] = p.edge_tab[16 + lane];
	__syncwarp();
	W.ideal = ideal;
	return W;
}
// row pointer (float2 or short2 samples) and its phase on the 16-byte grid
__device__ __forceinline__ const float2 *demod_row_f(const DemodParams &p, int b) { return reinterpret_cast<const float2 *>(p.bursts) + (size_t)b * p.stride; }
__device__ __forceinline__ const short2 *demod_row_s(const DemodParams &p, int b) { return reinterpret_cast<const short2 *>(p.iq) + (size_t)b * p.iq_stride; }
template <bool I16>
__device__ __forceinline__ unsigned demod_row_phase(const DemodParams &p, int b)
{
	if constexpr (I16) return (unsigned)((reinterpret_cast<uintptr_t>(demod_row_s(p, b)) >> 2) & 3u);
	else return (unsigned)((reinterpret_cast<uintptr_t>(demod_row_f(p, b)) >> 3) & 1u);
}
// staging call for one burst (see stage_async)
template <bool I16>
__device__ __forceinline__ void demod_stage(const DemodParams &p, int b, int off2, float2 *Ub, unsigned bar, int lane, float2 &patch, int &patch_idx)
{
	if constexpr (I16) stage_async16(demod_row_s(p, b), off2, Ub, bar, lane, patch, patch_idx);
	else stage_async(demod_row_f(p, b), off2, Ub, bar, lane, patch, patch_idx);
}

// Everything the demodulator warp does for burst b once its scalars are known: for a detected burst (rc > 0) the window
// U was staged by demod_stage() and completes on mbarrier `bar` (waited on here with `parity`); (patch, patch_idx) is
// the straddling sample stage_async() handed back.  Shared by demod_kernel and nb_fused_kernel.
template <bool I16>
__device__ __forceinline__ void demod_one(const DemodParams &p, const DemodWarp &W, int b, int rc, float2 amp, float toa, float2 *U,
					   float2 patch, int patch_idx, unsigned bar, unsigned parity, int lane)
{
	float *const ostage = W.ostage;
	float2 *const decs = W.decs, *const yv = W.yv, *const ytop = W.ytop, *const yx = W.yx;
	// pull path: the slot's power measurement (every slot that is not switched off, detected or not)
	bool want_energy = false;
	if constexpr (I16) {
		want_energy = p.type_raw[b] != 0;
		if (want_energy && rc <= 0) energy_terms16(demod_row_s(p, b), nullptr, 0, lane, p.pw + (size_t)b * 80);
	}
	if (rc <= 0) {
		// undetected burst: only the deferred clipping report is left to do (sigProcLib.cpp:1746-1764)
		if (p.fix_clip && rc == 0 && (!p.type || type_known(load_type(p.type, b, 0)))) {
			float2 v[20];
#pragma unroll
			for (int k = 0; k < 20; k++) {
				const int i = lane + 32 * k;
				v[k] = make_float2(0.0f, 0.0f);
				if (i < 625) {
					if constexpr (I16) {
						v[k] = cvt_s2(__ldg(reinterpret_cast<const unsigned *>(demod_row_s(p, b)) + i));
					} else {
						v[k] = __ldg(&demod_row_f(p, b)[i]);
					}
				}
			}
			float mx = 0.0f;
#pragma unroll
			for (int k = 0; k < 20; k++) mx = fmaxf(mx, fmaxf(fabsf(v[k].x), fabsf(v[k].y)));
#pragma unroll
			for (int o = 16; o; o >>= 1)
				mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
			if (lane == 0 && mx > 30000.0f) {
				p.rc[b] = -2;
				if (p.flags) p.flags[b] |= 4;
			}
		}
		return;
	}

	// ---- per-burst scalars ----
	const float ian = __frcp_rn(norm2(amp));
	const float2 s = make_float2(amp.x * ian, -amp.y * ian); // (complex)1.0 / amp (soft bits carry a 1e-4 tolerance)
	const BurstGeom bg = burst_geom<I16>(toa, demod_row_phase<I16>(p, b));
	const int whole = bg.whole, f = bg.f, e = bg.e; // e: 0..1 (float rows) / 0..3 (int16 rows)
	const bool edge = (rc == 5);

	// ---- the burst's window: straddling sample, then wait for the bulk copies ----
	if (patch_idx >= 0) {
		if constexpr (I16) reinterpret_cast<int *>(U)[patch_idx] = __float_as_int(patch.x);
		else U[patch_idx] = patch;
	}
	__syncwarp();
	mbar_wait(bar, parity);
	if constexpr (I16) {
		if (want_energy) {
			const bool in_win = bg.off2 <= 0 && 316 - bg.off2 < 4 * kSlots16;
			energy_terms16(demod_row_s(p, b), in_win ? reinterpret_cast<const short2 *>(U) : nullptr, bg.off2, lane,
				       p.pw + (size_t)b * 80);
		}
	}

	// ---- main pass (transposed FIR), shared by GMSK and EDGE: lane owns window samples 20*lane ..
	//      20*lane+19 and accumulates into outputs i = 5*lane - 8 + m, m = 0..12; sample j meets output m
	//      with tap u = j + 32 - 4m (0 <= u <= 35), coefficient ce[u] = comp0[f][e][u] ----
	// 8-PSK: the equaliser's last output (symbol 147) reads the decimated samples up to 149 - the last six are never used
	const int nout = edge ? 150 : p.n_gmsk_soft;
	{
		float2 acc[13];
#pragma unroll
		for (int m = 0; m < 13; m++) acc[m] = make_float2(0.0f, 0.0f);
		const float4 *xc4 = reinterpret_cast<const float4 *>(U) + 10 * lane;
		// int16 rows: the same 20 samples as ten 8-byte pairs, starting (e & 2) samples in
		const uint2 *xc2 = reinterpret_cast<const uint2 *>(reinterpret_cast<const short2 *>(U) + 20 * lane + (e & 2));
		const float *__restrict__ ce = c_tab.comp0[f][e & 1];
#pragma unroll
		for (int h = 0; h < 2; h++) {
			// samples j = 4k + 2h and 4k + 2h + 1, k = 0..4
			float2 xa[5], xb[5];
#pragma unroll
			for (int k = 0; k < 5; k++) {
				if constexpr (I16) {
					const uint2 v = xc2[2 * k + h];
					xa[k] = cvt_s2(v.x);
					xb[k] = cvt_s2(v.y);
				} else {
					const float4 v = xc4[2 * k + h];
					xa[k] = make_float2(v.x, v.y);
					xb[k] = make_float2(v.z, v.w);
				}
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
		// lanes 30, 31 lack their right-hand neighbours: outputs >= 150 are finished by the split pass below
		if (lane < 30) {
			if (!edge) {
				// soft value = Re(z_i * sum), z_i = (1/amp) * (-j)^i, i = 5*lane + a  (i mod 4 = (lane + a) mod 4)
				float zx = (lane & 1) ? s.y : s.x, zy = (lane & 1) ? -s.x : s.y;
				if (lane & 2) { zx = -zx; zy = -zy; }
#pragma unroll
				for (int a = 0; a < 5; a++) {
					ostage[5 * lane + a] = fmaf(zx, fin[a].x, -zy * fin[a].y);
					const float t = zx; // z *= -j
					zx = zy;
					zy = -t;
				}
			} else {
#pragma unroll
				for (int a = 0; a < 5; a++) decs[2 + 5 * lane + a] = cscale(fin[a], s);
			}
		}
	}
	// ---- outputs 150 .. nout-1 (EDGE, or GMSK callers asking for all 156): 4 lanes per output, 9 taps each ----
	if (nout > 150) {
		const int i = 150 + (lane >> 2), part = lane & 3;
		const float *__restrict__ c = p.comp + (size_t)f * 16 * 36;
		float2 d = make_float2(0.0f, 0.0f);
		if (i < nout) {
#pragma unroll
			for (int tt = 0; tt < 9; tt++) {
				const int t = 9 * part + tt;
				if (t < 35) {
					const float ct = __ldg(&c[t]);
					d = ffma2(win_get<I16>(U, 4 * i + t + e), make_float2(ct, ct), d);
				}
			}
		}
		d.x += __shfl_xor_sync(0xffffffffu, d.x, 1);
		d.y += __shfl_xor_sync(0xffffffffu, d.y, 1);
		d.x += __shfl_xor_sync(0xffffffffu, d.x, 2);
		d.y += __shfl_xor_sync(0xffffffffu, d.y, 2);
		if (i < nout && part == 0) {
			if (edge) decs[2 + i] = cscale(d, s);
			else ostage[i] = soft_out(i, d, s);
		}
	}
	// The pass above is the full composite over the zero-extended window.  Where the reference's intermediate
	// vectors are truncated the dropped terms are subtracted: leading outputs lose decimator taps k < kmin
	// (history / samples shifted in from below), trailing ones taps k > kmax (the delayed vector ends at sample
	// 624 before the shift: delayed samples q >= q0 = 640 + whole never reach the decimator).
	const int nlead = min(nout, (max(15, 15 + whole) + 3) >> 2);
	const int q0 = 640 + whole;
	const bool top_trunc = q0 <= 4 * (nout - 1) + 15;
	__syncwarp();
	// ---- leading outputs: evaluated again, directly, with the composite truncated to the surviving decimator taps
	//      (comp[f][kmin], built on the host): 4 lanes per output, 9 taps each.  (The first form of this correction
	//      evaluated the delayed samples the dropped taps would have read - 20 taps on every lane - and subtracted their
	//      contribution: 110 of the kernel's 560 instructions per burst for four or five outputs.) ----
	for (int base = 0; base < nlead; base += 8) {
		const int i = base + (lane >> 2), part = lane & 3;
		const int kmin = max(0, max(15 - 4 * i, 15 - 4 * i + whole));
		const int kmax = min(15, 639 + whole - 4 * i);
		const bool direct = i < nlead && kmin <= kmax && kmax == 15;
		float2 d = make_float2(0.0f, 0.0f);
		if (direct) {
			const float *__restrict__ c = p.comp + ((size_t)f * 16 + kmin) * 36 + 9 * part;
			float ct[9];
#pragma unroll
			for (int tt = 0; tt < 9; tt++) ct[tt] = (9 * part + tt < 35) ? __ldg(&c[tt]) : 0.0f;
#pragma unroll
			for (int tt = 0; tt < 9; tt++) d = ffma2(win_get<I16>(U, 4 * i + 9 * part + tt + e), make_float2(ct[tt], ct[tt]), d);
		}
		d.x += __shfl_xor_sync(0xffffffffu, d.x, 1);
		d.y += __shfl_xor_sync(0xffffffffu, d.y, 1);
		d.x += __shfl_xor_sync(0xffffffffu, d.x, 2);
		d.y += __shfl_xor_sync(0xffffffffu, d.y, 2);
		if (i < nlead && part == 0) {
			// no tap left at all: the reference's sample is an exact zero; a tail-truncated leading output (kmax < 15, a burst
			// hundreds of samples late) takes the generic two-stage path
			if (kmin <= kmax && kmax < 15) d = slow_output<I16>(U, 4 * i + e, f, kmin, kmax);
			if (edge) decs[2 + i] = cscale(d, s);
			else ostage[i] = soft_out(i, d, s);
		}
	}
	if (top_trunc) {
		// trailing outputs lose decimator taps k > kmax (the delayed vector ends at sample 624 before the shift: delayed samples
		// q >= q0 never reach the decimator): Y[q0 + m], m < 16, one per lane, then 4 lanes per output subtract g[k] * Y
		const int q0c = min(max(q0, 0), 644);
		float2 yt = make_float2(0.0f, 0.0f);
		if (lane < 16) {
			if (f < 64) {
#pragma unroll
				for (int j = 0; j < 20; j++) {
					const float hj = c_tab.delay[f][j];
					yt = ffma2(win_get<I16>(U, q0c + lane + e + j), make_float2(hj, hj), yt);
				}
			} else {
				yt = win_get<I16>(U, q0c + lane + e + 9);
			}
			ytop[lane] = yt;
		}
		__syncwarp();
		{
			// outputs whose taps reach q >= q0: at most 8 of them (Y beyond q0 + 8 is zero: the burst has ended)
			const int i = max(0, (q0 - 12) >> 2) + (lane >> 2), part = lane & 3;
			float2 d = make_float2(0.0f, 0.0f);
			if (i < nout) {
#pragma unroll
				for (int kk = 0; kk < 4; kk++) {
					const int m = 4 * i + 4 * part + kk - q0;
					if (m >= 0 && m < 16)
						d = ffma2(ytop[m], make_float2(W.gk[kk], W.gk[kk]), d);
				}
			}
			d.x += __shfl_xor_sync(0xffffffffu, d.x, 1);
			d.y += __shfl_xor_sync(0xffffffffu, d.y, 1);
			d.x += __shfl_xor_sync(0xffffffffu, d.x, 2);
			d.y += __shfl_xor_sync(0xffffffffu, d.y, 2);
			if (i < nout && i >= nlead && part == 0) {
				// no tap left at all (the reference's sample is an exact zero): store the zero, not a rounding residue
				const bool none = max(0, max(15 - 4 * i, 15 - 4 * i + whole)) > min(15, 639 + whole - 4 * i);
				if (edge) {
					const float2 c2 = cscale(d, s);
					decs[2 + i] = none ? make_float2(0.0f, 0.0f) : make_float2(decs[2 + i].x - c2.x, decs[2 + i].y - c2.y);
				} else {
					ostage[i] = none ? 0.0f : ostage[i] - soft_out(i, d, s);
				}
			}
		}
	}
	__syncwarp();
	if (edge) {
		demod_edge_tail(p, b, decs, lane, W.rot_lane, W.ideal);
		__syncwarp();
		return;
	}
	if constexpr (I16) {
		// ---- pull chain: the soft values leave as datagram bytes ----
		store_soft_bytes(p, b, ostage, nout, reinterpret_cast<uint8_t *>(yx), lane);
		__syncwarp();
		return;
	}
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
	__syncwarp();
}

// WPB warps per CTA, BPS CTAs per SM: 8 x 2, or 17 x 1 (one CTA of 17 warps fits the register file at 120 registers per
// thread and 223 KB of shared memory: one more resident warp per SM than two CTAs of 8)
template <bool I16, int WPB = 8, int BPS = 2>
__global__ void __launch_bounds__(WPB * 32, BPS)
demod_kernel(DemodParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	float2 *Ubase = reinterpret_cast<float2 *>(smem_raw) + (size_t)warp * (kDemodWarpFloats / 2);
	float *ostage = reinterpret_cast<float *>(Ubase + 2 * 2 * kBufSlots);
	const unsigned bar0 = smem_u32(ostage + kScratchFloats); // two 8-byte mbarriers, one per window buffer
	const int step = gridDim.x * wpb;
	const DemodWarp W = demod_warp_setup(p, ostage, lane);
	auto row_phase = [&](int b_) { return demod_row_phase<I16>(p, b_); };
	auto stage = [&](int b_, int off2, float2 *Ub, unsigned bar, float2 &patch, int &patch_idx) {
		demod_stage<I16>(p, b_, off2, Ub, bar, lane, patch, patch_idx);
	};

	if (lane == 0) {
		mbar_init(bar0, 1);
		mbar_init(bar0 + 8, 1);
	}
	fence_proxy_async();
	__syncwarp();

	int b = blockIdx.x * wpb + warp;
	// per-burst scalars run two bursts ahead of the filter, the staging copies one burst ahead
	int rc0 = 0, rc1 = 0;
	float2 amp0 = make_float2(1.0f, 0.0f), amp1 = amp0;
	float toa0 = 0.0f, toa1 = 0.0f;
	if (b < p.n) {
		rc0 = p.rc[b];
		amp0 = reinterpret_cast<const float2 *>(p.amp)[b];
		toa0 = p.toa[b];
	}
	if (b + step < p.n) {
		rc1 = p.rc[b + step];
		amp1 = reinterpret_cast<const float2 *>(p.amp)[b + step];
		toa1 = p.toa[b + step];
	}
	unsigned phase = 0; // bit k: parity the next wait on buffer k uses
	int cur = 0;
	float2 patch_n = make_float2(0.0f, 0.0f); // straddling sample of the burst being staged (lane 31)
	int patch_idx_n = -1;
	if (b < p.n && rc0 > 0) {
		const BurstGeom g0 = burst_geom<I16>(toa0, row_phase(b));
		stage(b, g0.off2, Ubase, bar0, patch_n, patch_idx_n);
	}
	for (; b < p.n; b += step, cur ^= 1) {
		const int rc = rc0;
		const float2 amp = amp0;
		const float toa = toa0;
		float2 *U = Ubase + (size_t)cur * 2 * kBufSlots;
		const float2 patch = patch_n;
		const int patch_idx = patch_idx_n;
		patch_idx_n = -1;
		// next burst: start its copies into the other buffer; burst after next: fetch its scalars
		{
			const int bn = b + step;
			rc0 = rc1; amp0 = amp1; toa0 = toa1;
			if (bn < p.n && rc0 > 0) {
				const BurstGeom gn = burst_geom<I16>(toa0, row_phase(bn));
				stage(bn, gn.off2, Ubase + (size_t)(cur ^ 1) * 2 * kBufSlots, bar0 + 8 * (cur ^ 1), patch_n, patch_idx_n);
			}
			const int bnn = bn + step;
			rc1 = 0;
			if (bnn < p.n) {
				rc1 = p.rc[bnn];
				amp1 = reinterpret_cast<const float2 *>(p.amp)[bnn];
				toa1 = p.toa[bnn];
			}
		}

		demod_one<I16>(p, W, b, rc, amp, toa, U, patch, patch_idx, bar0 + 8 * cur, (phase >> cur) & 1u, lane);
		if (rc > 0) phase ^= 1u << cur;
	}
}

// ---------------------------------------------------------------------------------------------
// demod1_kernel — demodAnyBurst at ONE sample per symbol (rx_sps = 1, sigProcLib.cpp:2030-2048 with sps == 1): the burst
// delayed by -toa (delay_vector_blk_kernel, the exact two-stage delayVector) is the 1-sps vector; what is left is the
// scaling by 1 / amp, GMSKReverseRotate + real part, or the 8-PSK tail shared with demod_kernel.  Outside every BASELINE
// configuration (those run at 4 sps): written for coverage, one warp per burst, not tuned.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
demod1_kernel(DemodParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	float *ostage = reinterpret_cast<float *>(smem_raw) + (size_t)warp * kScratchFloats;
	const DemodWarp W = demod_warp_setup(p, ostage, lane);
	const int len = p.sps1_len;
	for (int b = blockIdx.x * wpb + warp; b < p.n; b += gridDim.x * wpb) {
		const int rc = p.rc[b];
		if (rc <= 0) continue;
		const float2 amp = reinterpret_cast<const float2 *>(p.amp)[b];
		const float ian = __frcp_rn(norm2(amp));
		const float2 s = make_float2(amp.x * ian, -amp.y * ian);
		const float2 *row = reinterpret_cast<const float2 *>(p.bursts) + (size_t)b * p.stride;
		if (rc == 5) {
			for (int i = lane; i < len; i += 32) W.decs[2 + i] = cscale(row[i], s);
			__syncwarp();
			demod_edge_tail(p, b, W.decs, lane, W.rot_lane, W.ideal, len);
			__syncwarp();
		} else {
			const int nout = min(p.n_gmsk_soft, len);
			float *orow = p.soft + (size_t)b * p.soft_stride;
			for (int i = lane; i < nout; i += 32) orow[i] = soft_out(i, row[i], s);
		}
	}
}

} // namespace trxb200
