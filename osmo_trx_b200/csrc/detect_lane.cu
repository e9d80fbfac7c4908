// detect_lane.cu — detectAnyBurst for the common geometry (16-symbol sync sequence, max_toa <= 4: normal / EDGE /
// dummy bursts) with LANE = BURST: one kernel, no intermediates.
//
// corr_nb_kernel + peak_kernel (detect.cu) spread a group of 7 bursts over the lanes of a warp for the decimator and
// the correlator, write the correlation vectors and powers to global memory, and read them back lanes = bursts for the
// peak logic: two launches, 300 B of intermediates per burst, cross-lane staging, shuffles and 12 KB of shared memory
// per warp for the raw windows.  Here every thread runs the whole chain for its own burst:
//   decimate   35 outputs from the 152 samples of the correlator window, read straight from the thread's row in global
//              memory (a 32-byte sector serves four consecutive samples out of L1); a 16-sample register window slides
//              by four per output; sse_conv_real16 order on the packed pipe
//   correlate  20 outputs x 16 taps from the decimated samples in registers, sse_conv_cmplx_8n order, written to the
//              thread's column of the warp's [row][lane] tile
//   peak       peak_lane() of detect.cu on that column - the same function the two-kernel path runs
// Arithmetic per output is the two-kernel path's, bit for bit.  No lane idles (7-burst groups filled 28 or 31 of 32),
// nothing is staged, and the only shared memory is the correlation tile the peak logic indexes dynamically.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

constexpr int kDlWarps = 9;	 // warps per CTA, one CTA per SM
constexpr int kDlStagePitch = 17; // samples per row of the staging chunk (16 + 1: lanes = rows read conflict free)
struct DetLaneParams {
	CorrParams c;
	PeakParams q;
};
__host__ __device__ constexpr size_t det_lane_warp_bytes()
{
	return (size_t)(20 + 2 * kPadRows) * kRowPitch * sizeof(float2) + (size_t)35 * 32 * sizeof(float) + (size_t)32 * kDlStagePitch * sizeof(float2);
}
__host__ __device__ constexpr size_t det_lane_hdr_bytes() { return (size_t)kSinc512 * sizeof(float) + corr_nb_hdr_bytes(); }
__host__ __device__ constexpr size_t det_lane_smem() { return det_lane_hdr_bytes() + kDlWarps * det_lane_warp_bytes(); }

template <bool I16>
__global__ void __launch_bounds__(kDlWarps * 32, 1)
detect_lane_kernel(DetLaneParams P)
{
	extern __shared__ __align__(16) unsigned char dl_raw[];
	const CorrParams &cp = P.c;
	const PeakParams &p = P.q;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	float *stab = reinterpret_cast<float *>(dl_raw);
	float2 *hs = reinterpret_cast<float2 *>(dl_raw + (size_t)kSinc512 * sizeof(float)); // [SEQ_COUNT][kNbSeqPitch]
	SeqInfo *sinfo = reinterpret_cast<SeqInfo *>(dl_raw + (size_t)kSinc512 * sizeof(float) + (((size_t)SEQ_COUNT * kNbSeqPitch * 8 + 15) & ~(size_t)15));
	unsigned char *wb = dl_raw + det_lane_hdr_bytes() + (size_t)warp * det_lane_warp_bytes();
	float2 *C = reinterpret_cast<float2 *>(wb);					      // [kPadRows + 20 + kPadRows][32]
	float *Pw = reinterpret_cast<float *>(wb + (size_t)(20 + 2 * kPadRows) * kRowPitch * 8) + lane; // [35][32]
	const float2 *stg = reinterpret_cast<const float2 *>(wb + (size_t)(20 + 2 * kPadRows) * kRowPitch * 8 + (size_t)35 * 32 * 4); // [32][kDlStagePitch]
	const unsigned stg_s = (unsigned)__cvta_generic_to_shared(stg);
	for (int k = threadIdx.x; k < kSinc512; k += blockDim.x) stab[k] = p.sinc512[k];
	corr_nb_fill_hdr(hs, sinfo);
	for (int k = lane; k < (20 + 2 * kPadRows) * kRowPitch; k += 32) C[k] = make_float2(0.0f, 0.0f);
	__syncthreads();

	const float2 NZ = bc2(p.negzero);
	const float2 Z = make_float2(0.0f, 0.0f);
	float g16[16];
#pragma unroll
	for (int k = 0; k < 16; k++) g16[k] = c_tab.dnsamp[k];
	float2 *Cl0 = C + kPadRows * kRowPitch + lane; // row i of this lane's burst: Cl0[i * kRowPitch]

	const int ntiles = (p.n + 31) >> 5;
	for (int tile = blockIdx.x * kDlWarps + warp; tile < ntiles; tile += gridDim.x * kDlWarps) {
		const int b = tile * 32 + lane;
		const bool valid = b < p.n;
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
			const int sub = lane & 15, half = lane >> 4;
			auto issue = [&](int c) {
#pragma unroll
				for (int i = 0; i < 16; i++) {
					const int r = 2 * i + half;
					const unsigned long long rp = __shfl_sync(0xffffffffu, rowp, r);
					if (rp) {
						const unsigned dst = stg_s + SB * (unsigned)(r * kDlStagePitch + sub);
						const unsigned long long src = rp + SB * (unsigned)(16 * c + sub);
						if constexpr (I16) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
						else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
					}
				}
				asm volatile("cp.async.commit_group;" ::: "memory");
			};
			float2 X[28];
			issue(0);
#pragma unroll
			for (int c = 0; c < 10; c++) {
				asm volatile("cp.async.wait_group 0;" ::: "memory");
				__syncwarp();
#pragma unroll
				for (int t = 0; t < 16; t++) {
					if constexpr (I16) X[12 + t] = cvt_s2(reinterpret_cast<const unsigned *>(stg)[lane * kDlStagePitch + t]);
					else X[12 + t] = stg[lane * kDlStagePitch + t];
				}
				__syncwarp();
				if (c < 9) issue(c + 1);
				// outputs 4c - 3 .. 4c: output j reads window samples 4j .. 4j + 15 = X[4j - 16c + 12 ..]
#pragma unroll
				for (int o = 0; o < 4; o++) {
					const int j = 4 * c - 3 + o;
					if (j >= 0 && j < 35) {
						const int x0 = 4 * o; // 4j - 16c + 12
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
				for (int t = 0; t < 12; t++) X[t] = X[t + 16];
			}
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
	}
}

} // namespace trxb200
