// detect_lane.cu — detectAnyBurst for the common geometry (16-symbol sync sequence, max_toa <= 4: normal / EDGE /
// dummy bursts) with LANE = BURST: one kernel, no intermediates.
//
// corr_nb_kernel + peak_kernel (detect.cu) spread a group of 7 bursts over the lanes of a warp for the decimator and
// the correlator, write the correlation vectors and powers to global memory, and read them back lanes = bursts for the
// peak logic: two launches, 300 B of intermediates per burst, cross-lane staging, shuffles and 12 KB of shared memory
// per warp for the raw windows.  Here every thread runs the whole chain for its own burst:
//   windows    the 152-sample correlator windows (read as 160) arrive in sixteen-sample chunks: the warp copies the chunk
//              of all its 32 rows into shared memory with coalesced asynchronous copies, every lane takes its own row's
//              sixteen samples into registers, and the next chunk is in flight while four decimator outputs are evaluated
//   decimate   35 outputs, a 28-sample register window; sse_conv_real16 order on the packed pipe
//   correlate  20 outputs x 16 taps from the decimated samples in registers, sse_conv_cmplx_8n order, written to the
//              thread's column of the warp's [row][lane] tile
//   peak       peak_lane() of detect.cu on that column - the same function the two-kernel path runs
// Arithmetic per output is the two-kernel path's, bit for bit (the GPU parity suites run green on it), and the chain
// itself needs about 130 warp instructions per burst against 263.  The hard part is its input: 32 rows of 1,216 bytes,
// 5,000 bytes apart, per warp.  Measured per 2^20 bursts (corr_nb_kernel + peak_kernel: 0.53 ms): every thread loading its
// own row 1.58 ms (32 sector requests per load: lg_throttle); 8-byte cp.async chunks 0.68 ms (sixteen copies per chunk sit
// in the load/store queue until their data is back: mio_throttle); one TMA bulk copy per row and chunk 0.55 ms (the bulk
// copies are serialised through the uniform datapath: half of the kernel's instructions); TMA TILE copies over the rows
// taken two at a time (below) 0.42 ms - the default.  Rows that are not on the 16-byte grid, int16 rows and tiles with mixed
// window starts keep the per-row forms.  With the tile copies 36 % of the stall samples are waits for a chunk (ncu, profiles/
// r3a_detect_lane_summary.txt); asking for the rest of a tile's window ahead of time with four wide
// cp.async.bulk.prefetch.tensor boxes made it slower (0.475 against 0.44 ms): the limit is the rate at which the copy engine
// turns boxes of 144-byte rows into requests, not the latency of one of them.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

constexpr int kDlWarps = 12;	 // warps per CTA, one CTA per SM
constexpr int kDlChunk = 24;	  // float rows: window samples per chunk (seven chunks cover the 152-sample window)
constexpr int kDlChunk16 = 48;	  // int16 rows: window samples per chunk (four chunks; a box row of 52 I/Q pairs is as long as a float box row)
constexpr int kDlStagePitch = kDlChunk16 + 1; // int16 rows, cp.async form: words per row of the staging chunk (odd: lanes = rows read conflict free)
constexpr int kDlBox16 = kDlChunk16 + 4;	  // int16 rows, TMA form: words per box row (a chunk + the up to 3 samples in front of a window that starts off the 16-byte grid)
constexpr int kDlBulkPitch = kDlChunk + 2; // float rows, bulk copies: a chunk + the 2 samples a row that starts off the 16-byte grid needs (208 B)
struct DetLaneParams {
	const void *tmap; // CUtensorMap (in global memory, 64-byte aligned) over the burst rows taken two at a time (below); used when tma_on
	CorrParams c;
	PeakParams q;
	int tma_on; // float rows: window chunks arrive as TMA tiles
	int tma_shift; // 1: the burst array starts on an odd sample (8 bytes off the 16-byte grid, e.g. slots addressed in place in a
		       // resampled stream): the tensor starts one sample in front of it and every sample index moves up by one
	// Rounds after the first visit only the bursts the round before left undetected WITH a further attempt to make (EDGE ->
	// TSC fall-through): round r appends them to list_out (count in *list_n_out), round r + 1 walks list_in.  Without the
	// list a later round finds one or two such bursts in nearly every tile of 32 and pays a whole tile for each.
	const int *list_in, *list_n_in; // null: the tile's bursts are 32 consecutive rows
	int *list_out, *list_n_out;	 // null: no further round
};
__host__ __device__ constexpr size_t det_lane_warp_bytes()
{
	// correlation tile [kPadRows + 20 + kPadRows][32] + decimated powers [35][32] + two mbarriers, rounded to 1 KB (the
	// TMA swizzle pattern is a function of the shared address); the staging chunks of the decimator lie over the tile,
	// which is not in use while the windows are read
	return (((size_t)(20 + 2 * kPadRows) * kRowPitch * sizeof(float2) + (size_t)35 * 32 * sizeof(float) + 16) + 1023) & ~(size_t)1023;
}
__host__ __device__ constexpr size_t det_lane_hdr_bytes() { return ((size_t)kSinc512 * sizeof(float) + corr_nb_hdr_bytes() + 1023) & ~(size_t)1023; }
__host__ __device__ constexpr size_t det_lane_smem() { return det_lane_hdr_bytes() + kDlWarps * det_lane_warp_bytes(); }

// A row stride of 625 samples (5,000 bytes) is not a legal TMA stride, but TWO rows are (16 * stride bytes): the burst array
// is described to the TMA as [n / 2][4 * stride] floats.  One tile copy then brings the same 16 window samples of the 16 even
// rows of a warp's 32 bursts, a second one (inner coordinate + 2 * stride floats) those of the 16 odd rows: two instructions
// per chunk instead of one per row.  Boxes of 16 rows x 26 samples (208 bytes: the row pitch in shared memory skews the banks).
// The copy engine's cost is per box row rather than per byte: chunks of 24 samples (14 boxes per tile) instead of 16 (20 boxes).
// (32-sample chunks need 17 KB of staging per warp, i.e. ten warps instead of twelve: 0.42 against 0.40 ms.)
__device__ __forceinline__ void tma_load_2d(unsigned dst, const void *tmap, int c0, int c1, unsigned bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
		     "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
		     : "memory");
}

template <bool I16, bool LISTED = false>
__global__ void __launch_bounds__(kDlWarps * 32, 1)
detect_lane_kernel(const __grid_constant__ DetLaneParams P)
{
	extern __shared__ __align__(1024) unsigned char dl_raw[];
	const CorrParams &cp = P.c;
	const PeakParams &p = P.q;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	float *stab = reinterpret_cast<float *>(dl_raw);
	float2 *hs = reinterpret_cast<float2 *>(dl_raw + (size_t)kSinc512 * sizeof(float)); // [SEQ_COUNT][kNbSeqPitch]
	SeqInfo *sinfo = reinterpret_cast<SeqInfo *>(dl_raw + (size_t)kSinc512 * sizeof(float) + (((size_t)SEQ_COUNT * kNbSeqPitch * 8 + 15) & ~(size_t)15));
	unsigned char *wb = dl_raw + det_lane_hdr_bytes() + (size_t)warp * det_lane_warp_bytes();
	float2 *C = reinterpret_cast<float2 *>(wb);					      // [kPadRows + 20 + kPadRows][32]
	float *Pw = reinterpret_cast<float *>(wb + (size_t)(20 + 2 * kPadRows) * kRowPitch * 8) + lane; // [35][32]
	const float2 *stg = reinterpret_cast<const float2 *>(wb); // [2][32][kDlStagePitch], over the tile
	static_assert((size_t)2 * 32 * kDlBulkPitch * sizeof(float2) <= (size_t)(20 + 2 * kPadRows) * kRowPitch * sizeof(float2) + (size_t)35 * 32 * sizeof(float),
		      "staging chunks fit the tile and the powers behind it (neither is in use while the windows are read)");
	const unsigned bar_s = (unsigned)__cvta_generic_to_shared(wb + (size_t)(20 + 2 * kPadRows) * kRowPitch * 8 + (size_t)35 * 32 * 4);
	if (lane == 0) {
		mbar_init(bar_s, 1);
		mbar_init(bar_s + 8, 1);
	}
	unsigned phase = 0; // bit k: parity of the next wait on staging buffer k
	const unsigned stg_s = (unsigned)__cvta_generic_to_shared(stg);
	for (int k = threadIdx.x; k < kSinc512; k += blockDim.x) stab[k] = p.sinc512[k];
	corr_nb_fill_hdr(hs, sinfo);
	for (int k = lane; k < (20 + 2 * kPadRows) * kRowPitch; k += 32) C[k] = make_float2(0.0f, 0.0f);
	fence_proxy_async(); // the mbarriers are visible to the bulk copies
	__syncthreads();

	const float2 NZ = bc2(p.negzero);
	const float2 Z = make_float2(0.0f, 0.0f);
	float g16[16];
#pragma unroll
	for (int k = 0; k < 16; k++) g16[k] = c_tab.dnsamp[k];
	float2 *Cl0 = C + kPadRows * kRowPitch + lane; // row i of this lane's burst: Cl0[i * kRowPitch]

	constexpr bool listed = LISTED; // (a variant of its own: the indirection costs the int16 kernel its last free registers)
	const int n_eff = listed ? *P.list_n_in : p.n;
	const int ntiles = (n_eff + 31) >> 5;
	for (int tile = blockIdx.x * kDlWarps + warp; tile < ntiles; tile += gridDim.x * kDlWarps) {
		const bool valid = tile * 32 + lane < n_eff;
		const int b = listed ? (valid ? P.list_in[tile * 32 + lane] : 0) : tile * 32 + lane;
		int type = 0, tsc = 0, T = 0, rc = 0;
		if (valid) {
			type = load_type(p.type, b, 0);
			tsc = p.tsc[b];
			T = p.max_toa[b];
			if (p.round > 0) rc = p.rc[b];
		}
		Attempt at;
		bool run = valid && attempt_runs(type, tsc, T, p.max_toa_bound, 35, p.round, rc, sinfo, at);
		if (run && sinfo[at.seq].len != 16) run = false; // never in this geometry (launch_detect selects it by max_seq_len)
		// ---- the correlator windows (152 samples per burst, read as 160) come in sixteen-sample chunks: the warp copies
		//      the chunk of all its 32 rows into shared memory with coalesced asynchronous copies (a row's 128 bytes per
		//      half warp), every lane then takes its own row's sixteen samples into registers, and the next chunk is in
		//      flight while four decimator outputs are evaluated.  (Each thread reading its own row from global memory
		//      is 32 sector requests per load instruction: 61 % of all stall samples were lg_throttle in that form.) ----
		const int s_lo = run ? 4 * (at.start - 15) - 15 : 0;
		const unsigned long long rowp = run ? (I16 ? reinterpret_cast<unsigned long long>(reinterpret_cast<const unsigned *>(cp.iq) + (size_t)b * cp.iq_stride + s_lo)
							   : reinterpret_cast<unsigned long long>(reinterpret_cast<const float2 *>(cp.bursts) + (size_t)b * cp.stride + s_lo))
					      : 0ull;
		const bool any = __ballot_sync(0xffffffffu, run) != 0u;
		float2 dec[35];
		if (any) {
			constexpr unsigned SB = I16 ? 4u : 8u; // bytes per sample
			// Rows of one tile usually share the window start (one burst type): then row r's window is rp0 + r * rowbytes and
			// no lane has to ask another for its pointer (the shuffles of the general form were a third of all stall samples)
			const unsigned runmask = __ballot_sync(0xffffffffu, run);
			const int lead = __ffs(runmask) - 1;
			const int s_lo0 = __shfl_sync(0xffffffffu, s_lo, lead);
			const bool uniform = !listed && __ballot_sync(0xffffffffu, run && s_lo != s_lo0) == 0u;
			const unsigned long long rowbytes = (unsigned long long)SB * (unsigned long long)(I16 ? cp.iq_stride : cp.stride);
			const unsigned long long rp0 = __shfl_sync(0xffffffffu, rowp, lead) - (unsigned long long)lead * rowbytes; // row 0 of the tile
			// int16 rows without the TMA: 4-byte asynchronous copies, eight lanes per row (32 contiguous bytes), four rows per instruction
			auto issue = [&](int c) {
				const unsigned buf = stg_s + 4u * (unsigned)((c & 1) * 32 * kDlStagePitch);
				const int rsub = lane >> 3, s8 = lane & 7;
#pragma unroll
				for (int i = 0; i < 8; i++) {
					const int r = 4 * i + rsub;
					unsigned long long rp;
					if (uniform) rp = ((runmask >> r) & 1u) ? rp0 + (unsigned long long)r * rowbytes : 0ull;
					else rp = __shfl_sync(0xffffffffu, rowp, r);
					if (rp) {
						const unsigned dst = buf + 4u * (unsigned)(r * kDlStagePitch + s8);
						const unsigned long long src = rp + 4u * (unsigned)(kDlChunk16 * c + s8);
#pragma unroll
						for (int k = 0; k < kDlChunk16 / 8; k++)
							asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 32u * (unsigned)k), "l"(src + 32u * (unsigned)k) : "memory");
					}
				}
				asm volatile("cp.async.commit_group;" ::: "memory");
			};
			// float rows: ONE bulk copy (TMA) per row and chunk, issued by the row's own lane - 144 bytes from the 16-byte
			// grid at or below the window start (e = 0 or 1 samples in front), completion on the buffer's mbarrier.  The
			// 8-byte cp.async form needs sixteen instructions per chunk, all of which sit in the load/store queue until
			// their data is back (37 % of the stall samples were mio_throttle).
			const int e = (int)((rowp >> 3) & 1ull);
			const unsigned long long rowa = rowp & ~15ull;
			const unsigned nact = __popc(runmask);
			auto issue_bulk = [&](int c) {
				const unsigned bar = bar_s + 8u * (unsigned)(c & 1);
				fence_proxy_async(); // the buffer's earlier generic accesses are ordered before the asynchronous writes
				__syncwarp();
				if (lane == 0) mbar_arrive_expect_tx(bar, nact * (unsigned)(kDlBulkPitch * 8));
				__syncwarp();
				// (one lane walking the rows instead - a bulk copy is a uniform-datapath instruction, 32 lanes issuing one each are
				// serialised at about twenty instructions apiece - was measured slower: 0.65 against 0.55 ms per 2^20 bursts)
				if (run) {
					bulk_g2s(stg_s + (unsigned)(((c & 1) * 32 + lane) * kDlBulkPitch * 8), reinterpret_cast<const void *>(rowa + (unsigned long long)(8 * kDlChunk * c)),
						 (unsigned)(kDlBulkPitch * 8), bar);
				}
			};
			// float rows on the 16-byte grid, one window start for the whole tile, rows inside the TMA's view: two tile copies
			bool tma = false;
			if constexpr (!I16) tma = P.tma_on && uniform && (((p.n & 1) == 0) || tile * 32 + 32 <= (p.n & ~1));
			else tma = P.tma_on && uniform && (((p.n & 3) == 0) || tile * 32 + 32 <= (p.n & ~3));
			// (the TMA wants the start of a box on the 16-byte grid: a row's chunk starts at the even sample at or below the window
			// sample and is kDlBulkPitch samples long; e_even / e_odd = the sample in front for the even and the odd rows of the tile)
			const int s_t = s_lo0 + P.tma_shift; // window start as a sample index of the tensor's rows
			const int e_even = s_t & 1, e_odd = (cp.stride + s_t) & 1;
			auto issue_tma = [&](int c) {
				const unsigned bar = bar_s + 8u * (unsigned)(c & 1);
				fence_proxy_async();
				__syncwarp();
				if (lane == 0) {
					mbar_arrive_expect_tx(bar, 2u * 16u * (unsigned)(kDlBulkPitch * 8));
					const unsigned dst = stg_s + (unsigned)((c & 1) * 2 * 16 * kDlBulkPitch * 8);
					const int c1 = tile * 16;
					tma_load_2d(dst, P.tmap, 2 * (s_t + kDlChunk * c - e_even), c1, bar);
					tma_load_2d(dst + (unsigned)(16 * kDlBulkPitch * 8), P.tmap, 2 * (cp.stride + s_t + kDlChunk * c - e_odd), c1, bar);
				}
			};
			// the lane's row in a staged chunk: box lane & 1 (even / odd rows), row lane >> 1, kDlBulkPitch samples per row, its window
			// sample t at t + e; rows 208 bytes apart and the two boxes one sample out of step: conflict-free 8-byte reads
			const int trow = (lane & 1) * 16 * kDlBulkPitch + (lane >> 1) * kDlBulkPitch + ((lane & 1) ? e_odd : e_even);
			// int16 rows: 2,500 bytes apart, so FOUR rows make a legal TMA row (16 * iq_stride bytes): the slot array is described as
			// [n / 4][4 * iq_stride] 4-byte I/Q pairs, a chunk of the warp's 32 slots is four boxes of 8 rows x kDlBox16 pairs (row class
			// r = slot & 3 at inner coordinate r * iq_stride + ...), each starting on the 16-byte grid at or below its window sample
			const int s_t16 = s_lo0 + P.tma_shift;
			const int cls = lane & 3;
			const int e16 = (cls * cp.iq_stride + s_t16) & 3; // samples in front of this lane's window in its box row
			auto issue_tma16 = [&](int c) {
				const unsigned bar = bar_s + 8u * (unsigned)(c & 1);
				fence_proxy_async();
				__syncwarp();
				if (lane == 0) {
					mbar_arrive_expect_tx(bar, 4u * 8u * (unsigned)(kDlBox16 * 4));
					const unsigned dst = stg_s + (unsigned)((c & 1) * 4 * 8 * kDlBox16 * 4);
#pragma unroll
					for (int r = 0; r < 4; r++) {
						const int x0 = r * cp.iq_stride + s_t16 + kDlChunk16 * c;
						tma_load_2d(dst + (unsigned)(r * 8 * kDlBox16 * 4), P.tmap, x0 & ~3, tile * 8, bar);
					}
				}
			};
			const int trow16 = cls * 8 * kDlBox16 + (lane >> 2) * kDlBox16 + e16;
			// The decimator consumes a chunk eight samples at a time: sub-step g brings window samples 8g .. 8g + 7 and completes the
			// outputs 2g - 3 and 2g - 2 (output j reads window samples 4j .. 4j + 15), from a register window of 12 + 8 samples.
			constexpr int CS = I16 ? kDlChunk16 : kDlChunk, NCH = (152 + CS - 1) / CS, SUB = CS / 8;
			float2 X[20];
			__syncwarp(); // the previous tile's peak logic is done with the tile the chunks overlay
			if constexpr (I16) {
				if (tma) { issue_tma16(0); issue_tma16(1); }
				else { issue(0); issue(1); }
			} else if (tma) { issue_tma(0); issue_tma(1); }
			else { issue_bulk(0); issue_bulk(1); }
#pragma unroll
			for (int c = 0; c < NCH; c++) {
				const float2 *row = stg;
				const unsigned *row16 = reinterpret_cast<const unsigned *>(stg);
				if constexpr (I16) {
					if (tma) {
						mbar_wait(bar_s + 8u * (unsigned)(c & 1), (phase >> (c & 1)) & 1u);
						phase ^= 1u << (c & 1);
						row16 += (c & 1) * 4 * 8 * kDlBox16 + trow16;
					} else {
						if (c < NCH - 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
						else asm volatile("cp.async.wait_group 0;" ::: "memory");
						__syncwarp();
						row16 += ((c & 1) * 32 + lane) * kDlStagePitch;
					}
				} else {
					mbar_wait(bar_s + 8u * (unsigned)(c & 1), (phase >> (c & 1)) & 1u);
					phase ^= 1u << (c & 1);
					row = tma ? stg + (c & 1) * 2 * 16 * kDlBulkPitch + trow : stg + ((c & 1) * 32 + lane) * kDlBulkPitch + e;
				}
#pragma unroll
				for (int u = 0; u < SUB; u++) {
					if constexpr (I16) {
#pragma unroll
						for (int t = 0; t < 8; t++) X[12 + t] = cvt_s2(row16[8 * u + t]);
					} else {
#pragma unroll
						for (int t = 0; t < 8; t++) X[12 + t] = row[8 * u + t];
					}
					if (u == SUB - 1) {
						// every lane holds the rest of the chunk: its buffer takes the chunk after next
						__syncwarp();
						if (c + 2 < NCH) {
							if constexpr (I16) {
								if (tma) issue_tma16(c + 2);
								else issue(c + 2);
							} else if (tma) issue_tma(c + 2);
							else issue_bulk(c + 2);
						}
					}
					const int g = c * SUB + u;
#pragma unroll
					for (int o = 0; o < 2; o++) {
						const int j = 2 * g - 3 + o;
						if (j >= 0 && j < 35) {
							const int x0 = 4 * o; // window sample 4j = 8g - 12 + 4o
							float2 L[4];
#pragma unroll
							for (int q = 0; q < 4; q++) {
								const float2 p0 = mul2(X[x0 + q], bc2(g16[q]), NZ), p1 = mul2(X[x0 + 4 + q], bc2(g16[4 + q]), NZ);
								const float2 p2 = mul2(X[x0 + 8 + q], bc2(g16[8 + q]), NZ), p3 = mul2(X[x0 + 12 + q], bc2(g16[12 + q]), NZ);
								L[q] = add2(add2(p0, p1), add2(p2, p3));
							}
							dec[j] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
						}
					}
#pragma unroll
					for (int t = 0; t < 12; t++) X[t] = X[t + 8];
				}
			}
			// the chunks lay over the tile: its zero rows in front of and behind the correlation vector are restored (a lane
			// clears its own column of every pad row; the 20 vector rows are rewritten or cleared by the code below)
			__syncwarp();
#pragma unroll
			for (int r = 0; r < kPadRows; r++) {
				C[r * kRowPitch + lane] = make_float2(0.0f, 0.0f);
				C[(kPadRows + 20 + r) * kRowPitch + lane] = make_float2(0.0f, 0.0f);
			}
#pragma unroll
			for (int r = 0; r < 20; r++) Cl0[r * kRowPitch] = make_float2(0.0f, 0.0f);
			__syncwarp();
		}
		if (run) {
#pragma unroll
			for (int j = 0; j < 35; j++) Pw[j * 32] = norm2(dec[j]);
			// ---- correlation (sse_conv_cmplx_8n order, convolve_sse_3.c:462-537, h_len 16) ----
			const float2 *hh = hs + at.seq * kNbSeqPitch;
			float2 hr[16], hi[16];
#pragma unroll
			for (int k = 0; k < 16; k++) {
				const float2 h = hh[k];
				hr[k] = bc2(h.x);
				hi[k] = make_float2(h.y, -h.y);
			}
			const int len = at.len;
#pragma unroll
			for (int i = 0; i < 20; i++) {
				float2 L[4];
#pragma unroll
				for (int q = 0; q < 4; q++) {
					// A[q] and B[q] = the accumulators of taps (q, q + 8) and (q + 4, q + 12); L[q] = A[q] + B[q]
					const float2 a = add2(add2(Z, cmul_tap(dec[i + q], hr[q], hi[q], NZ)), cmul_tap(dec[i + q + 8], hr[q + 8], hi[q + 8], NZ));
					const float2 c = add2(add2(Z, cmul_tap(dec[i + q + 4], hr[q + 4], hi[q + 4], NZ)), cmul_tap(dec[i + q + 12], hr[q + 12], hi[q + 12], NZ));
					L[q] = add2(a, c);
				}
				const float2 out = add2(add2(L[0], L[1]), add2(L[2], L[3]));
				if (i < len) Cl0[i * kRowPitch] = out;
			}
		}
		const PeakRes res = peak_lane(p, sinfo, stab, Cl0, valid, run, at, type, tsc, T, rc, NZ, [&](int j) { return Pw[j * 32]; });
		if (valid) {
			if (res.write_all) {
				p.rc[b] = res.rc;
				reinterpret_cast<float2 *>(p.amp)[b] = res.amp;
				p.toa[b] = res.toa;
				p.ci[b] = res.ci;
				if (p.tsc_out) p.tsc_out[b] = (uint8_t)res.tsc_out;
			}
			if (p.flags) {
				if (p.round == 0) p.flags[b] = (uint8_t)res.flags;
				else if (res.flags) p.flags[b] |= (uint8_t)res.flags;
			}
		}
		if (P.list_out) {
			// every burst the next round can touch: still undetected and a further attempt exists (peak_lane's own conditions)
			bool again = false;
			if (valid && (res.write_all ? res.rc : rc) == 0 && type_known(type) && !((type == 1 || type == 5) && tsc > 7) && T <= p.max_toa_bound)
				again = make_attempt(type, tsc, T, p.round + 1).seq >= 0;
			const unsigned am = __ballot_sync(0xffffffffu, again);
			if (am) {
				int base = 0;
				if (lane == 0) base = atomicAdd(P.list_n_out, __popc(am));
				base = __shfl_sync(0xffffffffu, base, 0);
				if (again) P.list_out[base + __popc(am & ((1u << lane) - 1u))] = b;
			}
		}
	}
}

} // namespace trxb200
