// nbfused.cu — detectAnyBurst + demodAnyBurst for the common geometry (16-symbol sync sequence, max_toa <= 4, one
// detection attempt: normal / EDGE / dummy bursts) as ONE persistent, warp-specialised kernel for sm_100a.
//
// The three kernels of the general path (corr_nb_kernel -> peak_kernel -> demod_kernel) run back to back there: the
// correlators (FP32 / shared-memory work on 1.2 KB of each burst) and the serial peak search hide nothing of the
// HBM-bound demodulator and vice versa.  Here one CTA per SM keeps all three going at once on different bursts:
//
//   correlator warps (kFuNC)   groups of 7 bursts: window copies by cp.async one group ahead, decimate, correlate -
//        |                     corr_nb_run() of detect.cu, bit for bit - leaving the correlation vectors and the
//        |  tile_full/free     decimated powers in a shared-memory tile [sample][burst] (28 bursts = 4 groups, two
//        v                     tile slots) instead of the global intermediates
//   peak warp (1)              lanes = the tile's bursts: peak_lane() of detect.cu (gates, 9-step TOA bisection, C/I,
//        |                     amp); writes rc / amp / toa / ci / tsc / flags to global memory and (rc, amp, toa) into
//        |  res_full/free      a ring of kFuRing result slots
//        v
//   demodulator warps (kFuND)  one burst each, round robin over the CTA's burst sequence: scalars from the ring, the
//                              burst's window by TMA bulk copy one burst ahead (double buffered, per-warp mbarriers),
//                              demod_one() of demod.cu (composite FIR, edge corrections, EDGE tail, clip report)
//
// All hand-offs are mbarriers in shared memory (arrive = release, try_wait = acquire, CTA scope); every wait depends
// on strictly earlier tiles, so the pipeline cannot deadlock.  The correlator window of a burst (1,216 B) is read from
// HBM a few microseconds before the demodulator's bulk copy of the whole burst, which therefore finds those lines in
// L2: a burst crosses the HBM interface once.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

constexpr int kFuTile = 28;   // bursts per tile: 4 correlator groups of 7 = the lanes of one peak pass
constexpr int kFuRing = 4;    // result slots (tiles) between the peak warp and the demodulator warps
constexpr int kFuNC = 2;      // correlator warps
constexpr int kFuNP = 1;      // peak warps
constexpr int kFuND = 10;     // demodulator warps
constexpr int kFuThreads = 32 * (kFuNC + kFuNP + kFuND);
constexpr int kFuSlotRows = 20 + kPadRows; // tile slot s holds its 20 rows from row kPadRows + s * kFuSlotRows; pad rows are shared

struct FusedParams {
	CorrParams c;
	PeakParams q;
	DemodParams d;
};

// shared-memory layout (bytes, every part a multiple of 16)
constexpr size_t kFuOffSinc = 0;
constexpr size_t kFuOffHdr = kFuOffSinc + (size_t)kSinc512 * sizeof(float);
constexpr size_t kFuHdrBytes = (((size_t)SEQ_COUNT * kNbSeqPitch * 8 + 15) & ~(size_t)15) + ((SEQ_COUNT * sizeof(SeqInfo) + 15) & ~(size_t)15);
constexpr size_t kFuOffC = kFuOffHdr + kFuHdrBytes;
constexpr size_t kFuCRows = kPadRows + 2 * kFuSlotRows;
constexpr size_t kFuOffP = kFuOffC + kFuCRows * kRowPitch * sizeof(float2);
constexpr size_t kFuOffRing = kFuOffP + (size_t)2 * 35 * 32 * sizeof(float);
constexpr size_t kFuOffBar = kFuOffRing + (size_t)kFuRing * 32 * sizeof(float4);
constexpr int kFuBars = 2 + 2 + 2 * kFuRing; // tile_full[2], tile_free[2], res_full[kFuRing], res_free[kFuRing]
constexpr size_t kFuOffCorr = kFuOffBar + (((size_t)kFuBars * 8 + 15) & ~(size_t)15);
constexpr size_t kFuOffDemod = kFuOffCorr + (size_t)kFuNC * corr_nb_warp_bytes();
constexpr size_t kFuSmemBytes = kFuOffDemod + (size_t)kFuND * kDemodWarpFloats * sizeof(float);
static_assert(kFuSmemBytes <= 227 * 1024, "nb_fused_kernel: shared memory");
static_assert(corr_nb_warp_bytes() % 16 == 0 && (kDemodWarpFloats * sizeof(float)) % 16 == 0, "nb_fused_kernel: alignment");

// tile tl (CTA-local count) of this CTA is global tile blockIdx.x + tl * gridDim.x; group G = 4 * tl + gi
struct NbTileSched {
	int cw, ntiles;
	__device__ __forceinline__ int operator()(int q) const
	{
		const int G = cw + q * kFuNC;
		const long tile = (long)blockIdx.x + (long)(G >> 2) * gridDim.x;
		return tile < ntiles ? (int)tile * kFuTile + (G & 3) * kNbGroup : -1;
	}
};
struct NbTileSink {
	float2 *C;
	float *P;
	unsigned bar_full, bar_free;
	int cw;
	int s, col;
	__device__ __forceinline__ void begin(int q, int)
	{
		const int G = cw + q * kFuNC, tl = G >> 2;
		s = tl & 1;
		col = (G & 3) * kNbGroup;
		if (tl >= 2) mbar_wait(bar_free + 8u * s, (unsigned)(((tl >> 1) - 1) & 1)); // the peak warp is done with this slot's previous tile
	}
	__device__ __forceinline__ void pw(int, int g, int j, float v) { P[(s * 35 + j) * 32 + col + g] = v; }
	__device__ __forceinline__ void co(int, int g, int i, float2 v) { C[(kPadRows + s * kFuSlotRows + i) * kRowPitch + col + g] = v; }
	__device__ __forceinline__ void end(int, int lane)
	{
		__syncwarp();
		if (lane == 0) mbar_arrive(bar_full + 8u * s);
	}
};

__global__ void __launch_bounds__(kFuThreads, 1)
nb_fused_kernel(FusedParams fp)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	float *stab = reinterpret_cast<float *>(smem_raw + kFuOffSinc);
	float2 *hs = reinterpret_cast<float2 *>(smem_raw + kFuOffHdr);
	SeqInfo *sinfo = reinterpret_cast<SeqInfo *>(smem_raw + kFuOffHdr + (((size_t)SEQ_COUNT * kNbSeqPitch * 8 + 15) & ~(size_t)15));
	float2 *C = reinterpret_cast<float2 *>(smem_raw + kFuOffC);
	float *P = reinterpret_cast<float *>(smem_raw + kFuOffP);
	float4 *ring = reinterpret_cast<float4 *>(smem_raw + kFuOffRing);
	const unsigned bar_tile_full = smem_u32(smem_raw + kFuOffBar);
	const unsigned bar_tile_free = bar_tile_full + 16u;
	const unsigned bar_res_full = bar_tile_free + 16u;
	const unsigned bar_res_free = bar_res_full + 8u * kFuRing;
	const int n = fp.c.n;
	const int ntiles = (n + kFuTile - 1) / kFuTile;

	// ---- CTA set-up: tables, zeroed tile rows, barriers ----
	for (int k = threadIdx.x; k < kSinc512; k += blockDim.x) stab[k] = fp.q.sinc512[k];
	corr_nb_fill_hdr(hs, sinfo);
	for (int k = threadIdx.x; k < (int)kFuCRows * kRowPitch; k += blockDim.x) C[k] = make_float2(0.0f, 0.0f);
	if (threadIdx.x == 0) {
		for (int s = 0; s < 2; s++) {
			mbar_init(bar_tile_full + 8u * s, 4);  // one arrival per correlator group of the tile
			mbar_init(bar_tile_free + 8u * s, 1);  // the peak warp
		}
		for (int r = 0; r < kFuRing; r++) {
			mbar_init(bar_res_full + 8u * r, 1);	  // the peak warp
			mbar_init(bar_res_free + 8u * r, kFuND); // every demodulator warp, once per tile
		}
	}
	float *ostage = nullptr;
	float2 *Ubase = nullptr;
	unsigned bar0 = 0;
	if (warp >= kFuNC + kFuNP) {
		Ubase = reinterpret_cast<float2 *>(smem_raw + kFuOffDemod) + (size_t)(warp - kFuNC - kFuNP) * (kDemodWarpFloats / 2);
		ostage = reinterpret_cast<float *>(Ubase + 2 * 2 * kBufSlots);
		bar0 = smem_u32(ostage + kScratchFloats); // two 8-byte mbarriers, one per window buffer
		if (lane == 0) {
			mbar_init(bar0, 1);
			mbar_init(bar0 + 8, 1);
		}
	}
	fence_proxy_async();
	__syncthreads();

	if (warp < kFuNC) {
		// ================= correlator warps =================
		NbTileSched sched;
		sched.cw = warp; sched.ntiles = ntiles;
		NbTileSink sink;
		sink.C = C; sink.P = P; sink.bar_full = bar_tile_full; sink.bar_free = bar_tile_free; sink.cw = warp; sink.s = 0; sink.col = 0;
		corr_nb_run<false>(fp.c, hs, sinfo, smem_raw + kFuOffCorr + corr_nb_warp_bytes() * warp, lane, sched, sink);
	} else if (warp < kFuNC + kFuNP) {
		// ================= peak warp =================
		const PeakParams &p = fp.q;
		const float2 NZ = bc2(p.negzero);
		for (int tl = 0;; tl++) {
			const long tile = (long)blockIdx.x + (long)tl * gridDim.x;
			if (tile >= ntiles) break;
			const int s = tl & 1, r = tl % kFuRing;
			const int b = (int)tile * kFuTile + lane;
			const bool valid = lane < kFuTile && b < n;
			int type = 0, tsc = 0, T = 0;
			if (valid) { type = load_type(p.type, b, 0); tsc = p.tsc[b]; T = p.max_toa[b]; }
			Attempt at;
			const bool run = valid && attempt_runs(type, tsc, T, p.max_toa_bound, p.ndmax, 0, 0, sinfo, at);
			mbar_wait(bar_tile_full + 8u * s, (unsigned)((tl >> 1) & 1));
			const float *Ps = P + s * 35 * 32 + lane;
			const PeakRes res = peak_lane(p, sinfo, stab, C + (kPadRows + s * kFuSlotRows) * kRowPitch + lane, valid, run, at, type, tsc, T,
						      0, NZ, [&](int j) { return Ps[j * 32]; });
			// the ring slot's previous tile has been read by every demodulator warp
			if (tl >= kFuRing) mbar_wait(bar_res_free + 8u * r, (unsigned)((tl / kFuRing - 1) & 1));
			if (valid) {
				// round 0 of the only attempt: every output of a valid burst is defined here
				p.rc[b] = res.rc;
				reinterpret_cast<float2 *>(p.amp)[b] = res.amp;
				p.toa[b] = res.toa;
				p.ci[b] = res.ci;
				if (p.tsc_out) p.tsc_out[b] = (uint8_t)res.tsc_out;
				if (p.flags) p.flags[b] = (uint8_t)res.flags;
			}
			// bursts beyond the batch: a negative rc that no error code uses, so that the demodulator does nothing at all
			ring[r * 32 + lane] = make_float4(__int_as_float(valid ? res.rc : -100), res.amp.x, res.amp.y, res.toa);
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(bar_tile_free + 8u * s);
				mbar_arrive(bar_res_full + 8u * r);
			}
		}
	} else {
		// ================= demodulator warps =================
		const DemodParams &p = fp.d;
		const int dw = warp - kFuNC - kFuNP;
		const DemodWarp W = demod_warp_setup(p, ostage, lane);
		// burst j of the CTA's sequence: tile j / 28 (CTA-local), lane j % 28 of it.  Returns false past the last tile.
		auto fetch = [&](int j, int &b, int &rc, float2 &amp, float &toa) {
			const int tl = j / kFuTile, k = j - tl * kFuTile;
			const long tile = (long)blockIdx.x + (long)tl * gridDim.x;
			if (tile >= ntiles) return false;
			const int r = tl % kFuRing;
			mbar_wait(bar_res_full + 8u * r, (unsigned)((tl / kFuRing) & 1));
			const float4 v = ring[r * 32 + k];
			rc = __float_as_int(v.x);
			amp = make_float2(v.y, v.z);
			toa = v.w;
			b = (int)tile * kFuTile + k;
			if (k + kFuND >= kFuTile) {
				// this warp's last burst of the tile: the slot may be refilled once every warp has said so
				__syncwarp();
				if (lane == 0) mbar_arrive(bar_res_free + 8u * r);
			}
			return true;
		};
		int j = dw;
		int b0 = 0, rc0 = 0;
		float2 amp0 = make_float2(1.0f, 0.0f);
		float toa0 = 0.0f;
		bool have = fetch(j, b0, rc0, amp0, toa0);
		unsigned phase = 0; // bit k: parity the next wait on buffer k uses
		int cur = 0;
		float2 patch_n = make_float2(0.0f, 0.0f); // straddling sample of the burst being staged (lane 31)
		int patch_idx_n = -1;
		if (have && rc0 > 0) {
			const BurstGeom g0 = burst_geom<false>(toa0, demod_row_phase<false>(p, b0));
			demod_stage<false>(p, b0, g0.off2, Ubase, bar0, lane, patch_n, patch_idx_n);
		}
		while (have) {
			const int b = b0, rc = rc0;
			const float2 amp = amp0;
			const float toa = toa0;
			float2 *U = Ubase + (size_t)cur * 2 * kBufSlots;
			const float2 patch = patch_n;
			const int patch_idx = patch_idx_n;
			patch_idx_n = -1;
			// next burst of this warp: its scalars from the ring, its copies into the other buffer
			j += kFuND;
			have = fetch(j, b0, rc0, amp0, toa0);
			if (have && rc0 > 0) {
				const BurstGeom gn = burst_geom<false>(toa0, demod_row_phase<false>(p, b0));
				demod_stage<false>(p, b0, gn.off2, Ubase + (size_t)(cur ^ 1) * 2 * kBufSlots, bar0 + 8 * (cur ^ 1), lane, patch_n,
						   patch_idx_n);
			}
			if (rc != -100) demod_one<false>(p, W, b, rc, amp, toa, U, patch, patch_idx, bar0 + 8 * cur, (phase >> cur) & 1u, lane);
			if (rc > 0) phase ^= 1u << cur;
			cur ^= 1;
		}
	}
}

} // namespace trxb200
