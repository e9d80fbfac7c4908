// detect.cu — batched burst detection (detectAnyBurst, sigProcLib.cpp:1926-1957) for sm_100a.
//
// Two kernels per detection attempt, joined by a small L2-resident intermediate (per burst the
// correlation vector and the powers of the decimated samples, 300 B for a normal burst):
//
//   corr_nb_kernel / corr_long_kernel  (register-blocked; warp = 7 bursts / one burst, see below)
//            stage the part of each burst the correlator window needs (152 of 625 samples for a
//            normal burst) into shared memory with asynchronous copies, decimate 4 sps -> 1 sps
//            (downsampleBurst :1587-1601) and correlates against the sync sequence (:1674).  Every
//            output is an independent dot product evaluated in the reference's float32 order (SSE3
//            summation trees of convolve_sse_3.c, no FMA contraction) on the packed FP32 pipe: one
//            FMUL2/FADD2 (mul/add.rn.f32x2) handles the (re,im) pair, so results stay bit-identical.
//   peak_kernel  (warp = tile of 32 bursts, lanes = bursts)
//            the inherently serial, comparison-driven part: argmax (:1120-1139), edge gate (:1683),
//            peak-to-average threshold (:1541-1571,1689), the 9-step early/late TOA bisection over
//            sinc-interpolated points (:1100-1118,1141-1186), C/I (:1608-1639), amp = xcorr/gain,
//            TOA bookkeeping.  The tile's correlation vectors sit in shared memory [sample][lane]
//            (conflict free) between zero rows, which makes the 21-tap interpolation branch free:
//            a tap outside the reference's summation range multiplies a zero sample and adds +0.
//            The interpolation weight of tap d at grid position F/512 is sinc(pi*|d-10-F/512|), i.e.
//            one entry of the table-sinc folded onto the 1/512 grid (43 KB in shared memory, tap-major
//            with bit-reversed columns): one per-lane base pointer per step, immediate offsets per tap.
// Multi-attempt types (EDGE -> TSC fall-through :1933-1941, EXT_RACH's three sequences :1793-1800)
// run further rounds of the two kernels; only bursts still undetected take part.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

namespace {

constexpr float kClipThresh = 30000.0f; // CLIP_THRESH sigProcLib.cpp:49

// ---- packed FP32 (sm_100), exact: both halves are IEEE round-to-nearest and nothing may be contracted.
// ptxas (12.9) fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever -fmad says (it even folds
// fma(fma(a,b,-0),1,c)), which would change the rounding of the decision-bearing sums.  The product is
// therefore written as fma(a, b, nz) with nz = -0.0f passed in at RUN time: a*b + (-0) is the correctly
// rounded product including the sign of a zero result, and an FFMA2 cannot be merged into the FADD2 that
// consumes it.  (Checked in SASS: FFMA2 ... UR.F32 followed by FADD2.)
__device__ __forceinline__ float2 mul2(float2 a, float2 b, float2 nz)
{
	unsigned long long ra, rb, rc, rd;
	ra = *reinterpret_cast<unsigned long long *>(&a);
	rb = *reinterpret_cast<unsigned long long *>(&b);
	rc = *reinterpret_cast<unsigned long long *>(&nz);
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
	return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
	unsigned long long ra, rb, rd;
	ra = *reinterpret_cast<unsigned long long *>(&a);
	rb = *reinterpret_cast<unsigned long long *>(&b);
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
	return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

// int16 I/Q pair (one 32-bit word) -> complex float without the conversion unit: 0x4B000000 | u is the float 2^23 + u
// for u < 2^23, so with u = int16 ^ 0x8000 (= value + 32768) one subtraction of 2^23 + 32768 leaves the value, exactly.
// One LOP3 / PRMT+LOP3 per component and a packed add per sample, all on the full-rate pipes.
__device__ __forceinline__ float2 cvt_s2(unsigned word)
{
	const unsigned lo = (word & 0x0000ffffu) ^ 0x4b008000u;
	const unsigned hi = __byte_perm(word, 0x4b000000u, 0x7632) ^ 0x00008000u;
	return add2(make_float2(__uint_as_float(lo), __uint_as_float(hi)), make_float2(-8421376.0f, -8421376.0f));
}

// internal burst type of the SCH search (trxb200_detect_sch_batch); never accepted from a caller's type array
constexpr int kTypeSchFull = 7;

struct Attempt {
	int seq;   // sequence id or -1
	int head;  // search symbols before target
	int start; // correlation start index (target - head - 1)
	int len;   // head + tail
	int rc_hit;
};

// window parameters per burst type and attempt (sigProcLib.cpp:1782-1924)
__device__ __forceinline__ Attempt make_attempt(int type, int tsc, int T, int attempt)
{
	Attempt a;
	a.seq = -1; a.head = 0; a.start = 0; a.len = 0; a.rc_hit = 0;
	switch (type) {
	case 5: // EDGE, then TSC
		if (attempt == 0) { a.seq = SEQ_EDGE + tsc; a.head = 6; a.start = 82 - 6 - 1; a.len = 6 + 6 + T; a.rc_hit = 5; }
		else if (attempt == 1) { a.seq = SEQ_MIDAMBLE + tsc; a.head = 10; a.start = 82 - 10 - 1; a.len = 10 + 6 + T; a.rc_hit = 1; }
		break;
	case 1:
		if (attempt == 0) { a.seq = SEQ_MIDAMBLE + tsc; a.head = 10; a.start = 71; a.len = 16 + T; a.rc_hit = 1; }
		break;
	case 2: // EXT_RACH
		if (attempt < 3) { a.seq = SEQ_RACH + attempt; a.head = 8; a.start = 48 - 8 - 1; a.len = 16 + T; a.rc_hit = 2; }
		break;
	case 3:
		if (attempt == 0) { a.seq = SEQ_RACH; a.head = 8; a.start = 39; a.len = 16 + T; a.rc_hit = 3; }
		break;
	case 6:
		if (attempt == 0) { a.seq = SEQ_DUMMY; a.head = 10; a.start = 71; a.len = 16 + T; a.rc_hit = 6; }
		break;
	case kTypeSchFull: // detectSCHBurst, SCH_DETECT_FULL :1805-1861: target 106, head 105, tail 51; returns detectBurst's rc
		if (attempt == 0) { a.seq = SEQ_SCH; a.head = 105; a.start = 0; a.len = 156; a.rc_hit = 1; }
		break;
	default:
		break;
	}
	return a;
}

__device__ __forceinline__ bool type_known(int type) { return type == 1 || type == 2 || type == 3 || type == 5 || type == 6 || type == kTypeSchFull; }
// burst type as the kernels see it: the SCH search ignores the per-burst arrays; a caller's type array cannot select it
__device__ __forceinline__ int load_type(const uint8_t *type, int b, int sch)
{
	if (sch) return kTypeSchFull;
	const int t = type[b];
	return t == kTypeSchFull ? 0x7f : t;
}

// Does burst (type, tsc, T) run attempt `round`?  Shared by both kernels so that they agree.
// seq_len: length of the attempt's sync sequence (from the SeqInfo table the caller holds).
__device__ __forceinline__ bool attempt_runs(int type, int tsc, int T, int bound, int ndmax, int round, int rc_prev,
					      const SeqInfo *sinfo, Attempt &at)
{
	at = make_attempt(type, tsc, T, round);
	if (!type_known(type)) return false;
	if ((type == 1 || type == 5) && tsc > 7) return false;
	if (T > bound) return false;
	if (at.seq < 0) return false;
	if (sinfo[at.seq].len + at.len - 1 > ndmax) return false;
	if (round > 0 && rc_prev != 0) return false;
	return true;
}

__device__ __forceinline__ bool near_tie(float a, float b)
{
	const float m = fmaxf(fabsf(a), fabsf(b));
	return fabsf(a - b) <= 4.0f * 1.1920929e-7f * m;
}

} // namespace

// per-burst attempt parameters packed into one word: active | hlen << 1 | start << 8 | len << 16
__device__ __forceinline__ int pk_active(int w) { return w & 1; }
__device__ __forceinline__ int pk_hlen(int w) { return (w >> 1) & 127; }
__device__ __forceinline__ int pk_start(int w) { return (w >> 8) & 255; }
__device__ __forceinline__ int pk_len(int w) { return (w >> 16) & 0xffff; }

// ---------------------------------------------------------------------------------------------
// corr_nb_kernel — corr_kernel for the common configuration (16-symbol sync sequence, correlation length <= 20:
// normal / EDGE / dummy bursts at max_toa <= 4), register blocked so that the shared-memory data pipe, which
// bounds the generic kernel, carries each staged sample once per FOUR decimated outputs and each decimated
// sample once per FIVE correlation outputs.  Same arithmetic per output (order of convolve_sse_3.c), hence the
// same bits.
//   warp = group of 7 bursts.
//   stage     (cp.async, one group ahead of the arithmetic) 152 samples per burst as 76 16-byte slots; slot sl lives in plane sl & 7 at index sl >> 3 (plane
//             pitch 11, burst pitch 89 slots: both the lane-consecutive staging stores and the 8-slot-strided
//             decimation reads hit eight distinct 16-byte bank groups per quarter warp).
//   decimate  work item = (burst, quad a): outputs 4a .. 4a+3 from slots 8a .. 8a+13 (14 LDS.128); 63 items.
//   correlate work item = (burst, a): outputs 5a .. 5a+4 from decimated samples 5a .. 5a+19, which sit in five
//             planes (sample j in plane j % 5 at index j / 5), and the burst's sync sequence (rows skewed by 18
//             so that different sequences start on different banks); 28 items.
// ---------------------------------------------------------------------------------------------
constexpr int kNbGroup = 7;
constexpr int kNbSlots = 76;
constexpr int kNbPlanePitch = 11;
constexpr int kNbRawPitch = 89;
constexpr int kNbDecPitch = 36;
constexpr int kNbSeqPitch = 18;
__host__ __device__ constexpr size_t corr_nb_warp_bytes() { return (size_t)kNbGroup * kNbRawPitch * 16 + (size_t)kNbGroup * kNbDecPitch * 8; }
__host__ __device__ constexpr size_t corr_nb_hdr_bytes()
{
	return (((size_t)SEQ_COUNT * kNbSeqPitch * 8 + 15) & ~(size_t)15) + ((SEQ_COUNT * sizeof(SeqInfo) + 15) & ~(size_t)15);
}

__device__ __forceinline__ float2 cmul_tap(float2 xv, float2 hr, float2 hi, float2 NZ)
{
	const float2 p1 = mul2(xv, hr, NZ), p2 = mul2(xv, hi, NZ);
	return add2(p1, make_float2(p2.y, p2.x));
}

// ---- mbarrier primitives (shared-window 32-bit addresses); demod.cu adds the bulk-copy forms ----
__device__ __forceinline__ unsigned smem_u32(const void *ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
	unsigned done = 0;
	unsigned spins = 0;
	while (!done) {
		asm volatile("{\n\t.reg .pred p;\n\t"
			     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			     "selp.u32 %0, 1, 0, p;\n\t}"
			     : "=r"(done)
			     : "r"(bar), "r"(parity)
			     : "memory");
		if (!done && ++spins > (1u << 24)) __trap(); // a lost arrival must fail loudly, not hang the device
	}
}

// Where a correlator warp finds its groups and leaves its results.  Stand-alone kernel: groups strided over the grid,
// results to the global intermediates peak_kernel reads.  (nb_fused_kernel, nbfused.cu, supplies the tile forms.)
struct NbGridSched {
	int first, step, ngroups;
	// first burst of the warp's q-th group, or -1 when the warp has no q-th group
	__device__ __forceinline__ int operator()(int q) const
	{
		const long g = (long)first + (long)q * step;
		return g < ngroups ? (int)g * kNbGroup : -1;
	}
};
struct NbGlobalSink {
	float2 *corr; // [n][20]
	float *pwr;   // [n][35]
	__device__ __forceinline__ void begin(int, int) {}
	__device__ __forceinline__ void pw(int b0, int g, int j, float v) { pwr[(size_t)(b0 + g) * 35 + j] = v; }
	__device__ __forceinline__ void co(int b0, int g, int i, float2 v) { corr[(size_t)(b0 + g) * 20 + i] = v; }
	__device__ __forceinline__ void end(int, int) {}
};

// The correlator warp's whole life: hs / sinfo are the CTA's copies of the 16-symbol sequences and the sequence table
// (filled and synchronised by the caller), wbase the warp's staging area (corr_nb_warp_bytes()).
// I16: the windows are read from the radio's int16 slots (p.iq, pull chain): 8-byte slots, converted after the
// shared-memory read - detection then needs no float copy of the slot at all.
template <bool I16, class Sched, class Sink>
__device__ __forceinline__ void corr_nb_run(const CorrParams &p, const float2 *hs, const SeqInfo *sinfo, unsigned char *wbase, int lane,
					     Sched sched, Sink sink)
{
	float4 *raw = reinterpret_cast<float4 *>(wbase);			   // [kNbGroup][kNbRawPitch]
	float2 *dec = reinterpret_cast<float2 *>(raw + kNbGroup * kNbRawPitch); // [kNbGroup][kNbDecPitch]

	const float2 NZ = bc2(p.negzero);
	const float2 Z = make_float2(0.0f, 0.0f);
	float g16[16];
#pragma unroll
	for (int k = 0; k < 16; k++) g16[k] = c_tab.dnsamp[k];
	constexpr unsigned SB = I16 ? 8u : 16u; // bytes per slot (two samples)
	const unsigned base_par = I16 ? (unsigned)((reinterpret_cast<uintptr_t>(p.iq) >> 2) & 1u)
				      : (unsigned)((reinterpret_cast<uintptr_t>(p.bursts) >> 3) & 1u);
	const float2 *xall = reinterpret_cast<const float2 *>(p.bursts);
	const unsigned *qall = reinterpret_cast<const unsigned *>(p.iq); // one word per int16 sample
	const int rstride = I16 ? p.iq_stride : p.stride;

	// Software pipeline over the warp's groups: the window copies of the NEXT group (cp.async, global -> shared
	// without a register round trip) are issued as soon as the decimator has consumed the current windows and land
	// while the current group is correlated; the per-burst scalars (type, tsc, max_toa, rc) run one group further ahead.
	const unsigned raw_s = (unsigned)__cvta_generic_to_shared(raw) + SB * (unsigned)((lane & 7) * kNbPlanePitch + (lane >> 3));

	struct Scal { int type, tsc, T, rc; };
	auto load_scal = [&](int b0_) {
		Scal q;
		q.type = -1; q.tsc = 0; q.T = 0; q.rc = 0;
		const int b = b0_ + lane;
		if (lane < kNbGroup && b0_ >= 0 && b < p.n) {
			q.type = load_type(p.type, b, 0); q.tsc = p.tsc[b]; q.T = p.max_toa[b];
			if (p.round > 0) q.rc = p.rc[b];
		}
		return q;
	};
	// per-burst attempt parameters: active | start << 8 | len << 16 | seq << 24 (lanes 0..6)
	auto pack_scal = [&](const Scal &q) {
		int pk = 0;
		Attempt at;
		if (q.type >= 0 && attempt_runs(q.type, q.tsc, q.T, p.max_toa_bound, 35, p.round, q.rc, sinfo, at) && sinfo[at.seq].len == 16)
			pk = 1 | (at.start << 8) | (at.len << 16) | (at.seq << 24);
		return pk;
	};
	// slot sl = lane + 32 * it of burst g goes to plane sl & 7, index sl >> 3: per lane a fixed shared offset plus
	// immediates; row base, window start and alignment are warp-uniform per burst
	auto issue_copies = [&](int b0_, int pk_) {
#pragma unroll
		for (int g = 0; g < kNbGroup; g++) {
			const int w = __shfl_sync(0xffffffffu, pk_, g);
			const size_t row = (size_t)(b0_ + g) * (size_t)rstride;
			const int s_lo = 4 * (((w >> 8) & 255) - 15) - 15;
			const unsigned row_par = (base_par + (unsigned)(row & 1u)) & 1u;
			const bool aligned = (((unsigned)s_lo + row_par) & 1u) == 0;
			if (w & 1) {
#pragma unroll
				for (int it = 0; it < 3; it++) {
					if (lane + 32 * it < kNbSlots) {
						const unsigned dst = raw_s + SB * (unsigned)(g * kNbRawPitch + 4 * it);
						if constexpr (I16) {
							const unsigned *src = qall + row + s_lo + 2 * lane + 64 * it;
							if (aligned) {
								asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
							} else {
								asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
								asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u), "l"(src + 1) : "memory");
							}
						} else {
							const float2 *src = xall + row + s_lo + 2 * lane + 64 * it;
							if (aligned) {
								asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
							} else {
								asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
								asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u), "l"(src + 1) : "memory");
							}
						}
					}
				}
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};

	int q = 0;
	int b0 = sched(0);
	int my_pk = 0;
	Scal sc_next;
	if (b0 >= 0) {
		my_pk = pack_scal(load_scal(b0));
		issue_copies(b0, my_pk);
	}
	int b0_next = sched(1);
	sc_next = load_scal(b0_next);
	for (; b0 >= 0; q++) {
		const int pk_next = pack_scal(sc_next);
		const int b0_nn = b0_next >= 0 ? sched(q + 2) : -1;
		const bool any = __ballot_sync(0xffffffffu, my_pk & 1) != 0u;
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
		sink.begin(q, lane);
		if (!any) {
			// nothing to do in this group: keep the pipeline moving
			if (b0_next >= 0) issue_copies(b0_next, pk_next);
			sc_next = load_scal(b0_nn);
			sink.end(q, lane);
			my_pk = pk_next;
			b0 = b0_next; b0_next = b0_nn;
			continue;
		}

		// ---- decimation (sse_conv_real16 order, convolve_sse_3.c:188-264): 4 outputs per item ----
#pragma unroll
		for (int pass = 0; pass < 2; pass++) {
			const int it = lane + 32 * pass;
			const int g = min(it / 9, kNbGroup - 1);
			const int a = it - 9 * g;
			const int w = __shfl_sync(0xffffffffu, my_pk, g);
			if (it < 9 * kNbGroup && (w & 1)) {
				float4 s[14];
				if constexpr (I16) {
					const uint2 *r = reinterpret_cast<const uint2 *>(raw) + g * kNbRawPitch + a;
#pragma unroll
					for (int qq = 0; qq < 14; qq++) {
						const uint2 wv = r[(qq & 7) * kNbPlanePitch + (qq >> 3)];
						const float2 lo = cvt_s2(wv.x), hi = cvt_s2(wv.y);
						s[qq] = make_float4(lo.x, lo.y, hi.x, hi.y);
					}
				} else {
					const float4 *r = raw + g * kNbRawPitch + a;
#pragma unroll
					for (int qq = 0; qq < 14; qq++) s[qq] = r[(qq & 7) * kNbPlanePitch + (qq >> 3)];
				}
				float2 *dg = dec + g * kNbDecPitch;
#pragma unroll
				for (int o = 0; o < 4; o++) {
					float2 L[4];
#pragma unroll
					for (int qq = 0; qq < 4; qq++) {
						// taps q, 4+q, 8+q, 12+q: sample k of the output sits in slot 2o + k/2, half k & 1
						float2 pr[4];
#pragma unroll
						for (int m = 0; m < 4; m++) {
							const int k = 4 * m + qq;
							const float4 v = s[2 * o + (k >> 1)];
							pr[m] = mul2((k & 1) ? make_float2(v.z, v.w) : make_float2(v.x, v.y), bc2(g16[k]), NZ);
						}
						L[qq] = add2(add2(pr[0], pr[1]), add2(pr[2], pr[3]));
					}
					const float2 y = add2(add2(L[0], L[1]), add2(L[2], L[3]));
					const int j = 4 * a + o;
					if (j < 35) {
						const int j5 = (j * 13) >> 6; // j / 5 for j < 64
						dg[(j - 5 * j5) * 7 + j5] = y;
						sink.pw(b0, g, j, norm2(y));
					}
				}
			}
		}
		__syncwarp();
		// the windows are consumed: start the next group's copies and the scalars of the one after
		if (b0_next >= 0) issue_copies(b0_next, pk_next);
		sc_next = load_scal(b0_nn);

		// ---- correlation (sse_conv_cmplx_8n order, convolve_sse_3.c:462-537, h_len 16): 5 outputs per item ----
		{
			const int g = min(lane >> 2, kNbGroup - 1), a = lane & 3;
			const int w = __shfl_sync(0xffffffffu, my_pk, g);
			if (lane < 4 * kNbGroup && (w & 1)) {
				const float2 *dx = dec + g * kNbDecPitch + a;
				const float2 *hh = hs + ((w >> 24) & 255) * kNbSeqPitch;
				const int len = (w >> 16) & 255;
				float2 x[20];
#pragma unroll
				for (int t = 0; t < 20; t++) x[t] = dx[(t % 5) * 7 + t / 5];
				float2 A[5], L0[5], S01[5], out[5];
#pragma unroll
				for (int step = 0; step < 8; step++) {
					// tap pairs in the order (0,8) (4,12) (1,9) (5,13) (2,10) (6,14) (3,11) (7,15)
					const int qq = (step >> 1) + 4 * (step & 1);
					const float2 h1 = hh[qq], h2 = hh[qq + 8];
					const float2 h1r = bc2(h1.x), h1i = make_float2(h1.y, -h1.y);
					const float2 h2r = bc2(h2.x), h2i = make_float2(h2.y, -h2.y);
#pragma unroll
					for (int o = 0; o < 5; o++) {
						const float2 acc = add2(add2(Z, cmul_tap(x[o + qq], h1r, h1i, NZ)), cmul_tap(x[o + qq + 8], h2r, h2i, NZ));
						if ((step & 1) == 0) {
							A[o] = acc; // A[q]
						} else {
							const float2 Lq = add2(A[o], acc); // L[q - 4] = A[q - 4] + B[q - 4]
							if (step == 1) L0[o] = Lq;
							else if (step == 3) S01[o] = add2(L0[o], Lq);
							else if (step == 5) L0[o] = Lq;
							else out[o] = add2(S01[o], add2(L0[o], Lq));
						}
					}
				}
#pragma unroll
				for (int o = 0; o < 5; o++)
					if (5 * a + o < len) sink.co(b0, g, 5 * a + o, out[o]);
			}
		}
		sink.end(q, lane);
		my_pk = pk_next;
		b0 = b0_next; b0_next = b0_nn;
	}
}

// the CTA's copies of the 16-symbol sequences and the sequence table, in front of the warps' staging areas
__device__ __forceinline__ void corr_nb_fill_hdr(float2 *hs, SeqInfo *sinfo)
{
	for (int k = threadIdx.x; k < SEQ_COUNT * 16; k += blockDim.x) {
		const int id = k >> 4, t = k & 15;
		if (c_tab.info[id].len == 16) hs[id * kNbSeqPitch + t] = c_tab.seq[c_tab.info[id].off + t];
	}
	for (int k = threadIdx.x; k < SEQ_COUNT; k += blockDim.x) sinfo[k] = c_tab.info[k];
}

// WPB warps per CTA, BPS CTAs per SM: 8 x 2, or one CTA of 18 warps (216 KB of staging, 112 registers per thread)
template <bool I16, int WPB = 8, int BPS = 2>
__global__ void __launch_bounds__(WPB * 32, BPS)
corr_nb_kernel(CorrParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	float2 *hs = reinterpret_cast<float2 *>(smem_raw); // [SEQ_COUNT][kNbSeqPitch]
	SeqInfo *sinfo = reinterpret_cast<SeqInfo *>(smem_raw + (((size_t)SEQ_COUNT * kNbSeqPitch * 8 + 15) & ~(size_t)15));
	corr_nb_fill_hdr(hs, sinfo);
	__syncthreads();
	NbGridSched sched;
	sched.first = blockIdx.x * wpb + warp; sched.step = gridDim.x * wpb; sched.ngroups = (p.n + kNbGroup - 1) / kNbGroup;
	NbGlobalSink sink;
	sink.corr = p.corr; sink.pwr = p.pwr;
	corr_nb_run<I16>(p, hs, sinfo, smem_raw + corr_nb_hdr_bytes() + corr_nb_warp_bytes() * warp, lane, sched, sink);
}

// ---------------------------------------------------------------------------------------------
// corr_long_kernel — every other configuration (access bursts with their 40-symbol sync sequence and 60+ symbol
// search window, normal bursts with a wide max_toa): one warp per burst, register blocked like corr_nb_kernel.
// The arithmetic is bound by the FP32 pipe here (a RACH burst costs 79 x 40 complex taps = 12.6 k packed
// operations against 3.9 KB of samples), so the layout is chosen to keep everything else off the issue slots:
//   stage     cp.async copies of the NEXT burst's window (zero filled outside samples 0..623 through the
//             src-size operand) land in the second buffer while the current burst is processed.  Slot sl
//             (2 samples) lives in plane sl & 7 at index sl >> 3, plane pitch odd.
//   decimate  lane a: outputs 4a .. 4a+3 from slots 8a .. 8a+13 (14 conflict-free LDS.128).
//   correlate lane a: outputs 3a .. 3a+2; per block of 8 taps 10 decimated samples are loaded once and feed all
//             three outputs; the taps come from __constant__ (the sequence is the same for the whole warp);
//             24 independent accumulation chains per lane (A[q], B[q] of convolve_sse_3.c:462-537 per output).
// Same operations in the same order per output as corr_kernel / the reference, hence the same bits.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int corr_lg_pp(int ndmax) { return (((ndmax + 3) >> 2) + 2) | 1; } // plane pitch in slots
__host__ __device__ constexpr size_t corr_lg_warp_bytes(int ndmax)
{
	return (size_t)2 * 8 * corr_lg_pp(ndmax) * 16 + ((((size_t)ndmax + 8) * 8 + 15) & ~(size_t)15);
}

__host__ __device__ constexpr size_t corr_lg_hdr_bytes() { return (size_t)SEQ_STORE * sizeof(uint4); } // the CTA's copy of CorrParams::seq_pm1

// One tap of a rotated GMSK sequence: one of its components is +-1 exactly, the other a rounding residue of the rotator
// (6e-17 .. 4e-15).  x * (+-1) is a sign flip - exact, on the ALU pipe - and only the product with the residue stays on the
// FP32 pipe, which bounds this kernel: three packed operations per tap and output instead of four, same bits.
//   even tap (REAL): (x.re * hr, x.im * hr) = x ^ (m, m);  T = (x.re * hi, x.im * -hi);  result = E + swap(T)
//   odd tap (!REAL): T = (x.re * hr, x.im * hr);  (x.re * hi, x.im * -hi) = x ^ (m, ~m) = E;  result = T + swap(E)
template <bool REAL>
__device__ __forceinline__ float2 tap_pm1(float2 x, uint4 t, float2 NZ)
{
	const float2 E = make_float2(__uint_as_float(__float_as_uint(x.x) ^ t.x), __uint_as_float(__float_as_uint(x.y) ^ t.y));
	const float2 T = mul2(x, make_float2(__uint_as_float(t.z), __uint_as_float(t.w)), NZ);
	if (REAL) return add2(E, make_float2(T.y, T.x));
	return add2(T, make_float2(E.y, E.x));
}

// corr_long_items for the +-1 sequences.  Lane = (group of five outputs, half of the accumulation chains): the eight chains
// of an output (A[q], B[q], convolve_sse_3.c:462-537) are split q = 0, 1 / q = 2, 3 over the lanes l and l + 16, whose
// partial sums (L0 + L1), (L2 + L3) meet in one shuffle - the reference's order - so that an access burst's 79 outputs
// occupy all 32 lanes (16 groups x 2 halves) instead of 27 (x 3 outputs).  tp: the sequence's taps in shared memory.
template <int HLEN>
__device__ __forceinline__ void corr_long_items_pm1(const float2 *dec, const uint4 *tp, int len, int hlen_rt, int lane, float2 *crow, float2 NZ)
{
	const int hlen = HLEN ? HLEN : hlen_rt;
	// (half = lane >> 4: the sixteen lanes of a half warp read windows 5 samples apart - 5 is coprime to the 16 bank pairs, so the
	// 8-byte loads are conflict free; with the halves interleaved, lane = 2 g + half, a third of the wavefronts were conflicts)
	const int hf = lane >> 4;
	for (int g0 = 0; 5 * g0 < len; g0 += 16) {
		const int g = g0 + (lane & 15);
		const bool act = 5 * g < len;
		const float2 *dx = dec + (act ? 5 * g : 0) + 2 * hf;
		const uint4 *th = tp + 2 * hf;
		float2 A[2][5], B[2][5];
#pragma unroll
		for (int qq = 0; qq < 2; qq++)
#pragma unroll
			for (int o = 0; o < 5; o++) { A[qq][o] = make_float2(0.0f, 0.0f); B[qq][o] = make_float2(0.0f, 0.0f); }
#pragma unroll
		for (int t0 = 0; t0 < hlen; t0 += 8) {
			asm volatile("" ::: "memory"); // keeps the unrolled blocks' loads in their own block (register pressure)
			float2 xw[10];
#pragma unroll
			for (int k = 0; k < 10; k++) xw[k] = dx[t0 + k];
			const uint4 ta0 = th[t0], ta1 = th[t0 + 1], tb0 = th[t0 + 4], tb1 = th[t0 + 5];
#pragma unroll
			for (int o = 0; o < 5; o++) {
				A[0][o] = add2(A[0][o], tap_pm1<true>(xw[o], ta0, NZ));
				B[0][o] = add2(B[0][o], tap_pm1<true>(xw[4 + o], tb0, NZ));
				A[1][o] = add2(A[1][o], tap_pm1<false>(xw[1 + o], ta1, NZ));
				B[1][o] = add2(B[1][o], tap_pm1<false>(xw[5 + o], tb1, NZ));
			}
		}
#pragma unroll
		for (int o = 0; o < 5; o++) {
			const float2 P = add2(add2(A[0][o], B[0][o]), add2(A[1][o], B[1][o])); // L[2 hf] + L[2 hf + 1]
			float2 Q;
			Q.x = __shfl_xor_sync(0xffffffffu, P.x, 16);
			Q.y = __shfl_xor_sync(0xffffffffu, P.y, 16);
			if (act && hf == 0 && 5 * g + o < len) crow[5 * g + o] = add2(P, Q); // (L0 + L1) + (L2 + L3)
		}
	}
}

// HLEN > 0: compile-time sequence length (fully unrolled, immediate offsets); 0: run-time multiple of 8
template <int HLEN>
__device__ __forceinline__ void corr_long_items(const float2 *dec, const float2 *hh, int len, int hlen_rt, int lane, float2 *crow,
						 float2 NZ)
{
	const int hlen = HLEN ? HLEN : hlen_rt;
	for (int a = lane; 3 * a < len; a += 32) {
		const float2 *dx = dec + 3 * a;
		float2 A[4][3], B[4][3];
#pragma unroll
		for (int q = 0; q < 4; q++)
#pragma unroll
			for (int o = 0; o < 3; o++) { A[q][o] = make_float2(0.0f, 0.0f); B[q][o] = make_float2(0.0f, 0.0f); }
#pragma unroll
		for (int t0 = 0; t0 < hlen; t0 += 8) {
			asm volatile("" ::: "memory"); // keeps the unrolled blocks' loads in their own block (register pressure)
			float2 xw[10];
#pragma unroll
			for (int k = 0; k < 10; k++) xw[k] = dx[t0 + k];
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const float2 h1 = hh[t0 + q], h2 = hh[t0 + 4 + q];
				const float2 h1r = bc2(h1.x), h1i = make_float2(h1.y, -h1.y);
				const float2 h2r = bc2(h2.x), h2i = make_float2(h2.y, -h2.y);
#pragma unroll
				for (int o = 0; o < 3; o++) {
					A[q][o] = add2(A[q][o], cmul_tap(xw[q + o], h1r, h1i, NZ));
					B[q][o] = add2(B[q][o], cmul_tap(xw[4 + q + o], h2r, h2i, NZ));
				}
			}
		}
		float2 *co = crow + 3 * a;
#pragma unroll
		for (int o = 0; o < 3; o++) {
			float2 L[4];
#pragma unroll
			for (int q = 0; q < 4; q++) L[q] = add2(A[q][o], B[q][o]);
			if (3 * a + o < len) co[o] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
		}
	}
}

__global__ void __launch_bounds__(320, 2)
corr_long_kernel(CorrParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	const int ndmax = p.ndmax, lmax = p.lmax;
	const int PP = corr_lg_pp(ndmax);
	// the +-1 taps of every sequence in front of the warps' areas
	uint4 *pm1_s = p.seq_pm1 ? reinterpret_cast<uint4 *>(smem_raw) : nullptr;
	if (pm1_s) {
		for (int k = threadIdx.x; k < SEQ_STORE; k += blockDim.x) pm1_s[k] = p.seq_pm1[k];
		__syncthreads();
	}
	unsigned char *wbase = smem_raw + corr_lg_hdr_bytes() + corr_lg_warp_bytes(ndmax) * warp;
	float4 *raw0 = reinterpret_cast<float4 *>(wbase);		      // [2][8][PP]
	float2 *dec = reinterpret_cast<float2 *>(raw0 + (size_t)2 * 8 * PP); // [ndmax + 8]
	const unsigned raw_s = (unsigned)__cvta_generic_to_shared(raw0);

	const float2 NZ = bc2(p.negzero);
	float g16[16];
#pragma unroll
	for (int k = 0; k < 16; k++) g16[k] = c_tab.dnsamp[k];
	const unsigned base_par = (unsigned)((reinterpret_cast<uintptr_t>(p.bursts) >> 3) & 1u);
	const float2 *xall = reinterpret_cast<const float2 *>(p.bursts);
	const int step = gridDim.x * wpb;

	struct Scal { int type, tsc, T, rc; };
	auto load_scal = [&](int b_) {
		Scal q;
		q.type = -1; q.tsc = 0; q.T = 0; q.rc = 0;
		if (b_ < p.n) {
			q.type = load_type(p.type, b_, p.sch);
			if (!p.sch) { q.tsc = p.tsc[b_]; q.T = p.max_toa[b_]; }
			if (p.round > 0) q.rc = p.rc[b_];
		}
		return q;
	};
	// attempt parameters: x = active | hlen << 1 | start << 8 | len << 16, y = offset of the sequence in c_tab.seq
	auto plan = [&](const Scal &q) {
		int2 r = make_int2(0, 0);
		Attempt at;
		if (q.type >= 0 && attempt_runs(q.type, q.tsc, q.T, p.max_toa_bound, ndmax, p.round, q.rc, c_tab.info, at)) {
			r.x = 1 | (c_tab.info[at.seq].len << 1) | (at.start << 8) | (at.len << 16);
			r.y = c_tab.info[at.seq].off | (c_tab.info[at.seq].pm1 << 30);
		}
		return r;
	};
	auto issue_copies = [&](int b_, int pk_, int buf) {
		if (pk_active(pk_)) {
			const int nd = pk_hlen(pk_) + pk_len(pk_) - 1;
			const int s_lo = 4 * (pk_start(pk_) - pk_hlen(pk_) + 1) - 15;
			const int ns = 2 * nd + 6;
			const size_t row = (size_t)b_ * (size_t)p.stride;
			const float2 *x = xall + row;
			const unsigned row_par = (base_par + (unsigned)(row & 1u)) & 1u;
			const bool aligned = (((unsigned)s_lo + row_par) & 1u) == 0; // window slots sit on the row's 16-byte grid
			const unsigned dst0 = raw_s + 16u * (unsigned)(buf * 8 * PP);
			// slot sl = lane + 32 * it: plane sl & 7 = lane & 7, index sl >> 3 = (lane >> 3) + 4 * it
			unsigned dst = dst0 + 16u * (unsigned)((lane & 7) * PP + (lane >> 3));
			const float2 *src = x + s_lo + 2 * lane;
			int idx = s_lo + 2 * lane;
			for (int sl = lane; sl < ns; sl += 32, dst += 64u, src += 64, idx += 64) {
				if (idx >= 0 && idx <= 622) {
					if (aligned) {
						asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
					} else {
						asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
						asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u), "l"(src + 1) : "memory");
					}
				} else {
					// downsampleBurst reads samples 0..623 behind 16 zero history samples (:1590-1593): zero fill
					// through the src-size operand (the clamped address is never read when the size is 0)
					const unsigned n0 = (idx >= 0 && idx <= 623) ? 8u : 0u, n1 = (idx + 1 >= 0 && idx + 1 <= 623) ? 8u : 0u;
					const float2 *s0 = x + min(max(idx, 0), 623), *s1 = x + min(max(idx + 1, 0), 623);
					asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(s0), "r"(n0) : "memory");
					asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + 8u), "l"(s1), "r"(n1) : "memory");
				}
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};

	int b = blockIdx.x * wpb + warp;
	int2 pk0 = plan(load_scal(b));
	const bool sps1 = p.sps1_len > 0;
	if (b < p.n && !sps1) issue_copies(b, pk0.x, 0);
	Scal s1 = load_scal(b + step);
	int cur = 0;
	for (; b < p.n; b += step, cur ^= 1) {
		const int2 pk1 = plan(s1);
		s1 = load_scal(b + 2 * step);
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
		// the other buffer was consumed by the previous burst's decimation: refill it with the next window
		if (b + step < p.n && !sps1) issue_copies(b + step, pk1.x, cur ^ 1);
		if (pk_active(pk0.x)) {
			const int hlen = pk_hlen(pk0.x), len = pk_len(pk0.x);
			const int nd = hlen + len - 1;
			const int dstart = pk_start(pk0.x) - (hlen - 1);
			const float4 *raw = raw0 + (size_t)cur * 8 * PP;
			if (sps1) {
				// one sample per symbol: the "decimated" burst is the burst (zeros outside it: convolve()'s head-room, :312-334)
				const float2 *x = xall + (size_t)b * (size_t)p.stride;
				float *pw = p.pwr + (size_t)b * ndmax;
				for (int j = lane; j < nd; j += 32) {
					const int d = dstart + j;
					const float2 y = (d >= 0 && d < p.sps1_len) ? __ldg(&x[d]) : make_float2(0.0f, 0.0f);
					dec[j] = y;
					pw[j] = norm2(y);
				}
			}
			// ---- decimation (sse_conv_real16 order, convolve_sse_3.c:188-264): 4 outputs per item ----
			for (int a = lane; !sps1 && 4 * a < nd; a += 32) {
				float4 s[14];
#pragma unroll
				for (int q = 0; q < 14; q++) s[q] = raw[(q & 7) * PP + a + (q >> 3)];
				float *pw = p.pwr + (size_t)b * ndmax;
#pragma unroll
				for (int o = 0; o < 4; o++) {
					float2 L[4];
#pragma unroll
					for (int q = 0; q < 4; q++) {
						// taps q, 4+q, 8+q, 12+q: sample k of the output sits in slot 2o + k/2, half k & 1
						float2 pr[4];
#pragma unroll
						for (int m = 0; m < 4; m++) {
							const int k = 4 * m + q;
							const float4 v = s[2 * o + (k >> 1)];
							pr[m] = mul2((k & 1) ? make_float2(v.z, v.w) : make_float2(v.x, v.y), bc2(g16[k]), NZ);
						}
						L[q] = add2(add2(pr[0], pr[1]), add2(pr[2], pr[3]));
					}
					float2 y = add2(add2(L[0], L[1]), add2(L[2], L[3]));
					const int j = 4 * a + o, d = dstart + j;
					if (j < nd) {
						if (d < 0 || d >= 156) y = make_float2(0.0f, 0.0f);
						dec[j] = y;
						pw[j] = norm2(y);
					}
				}
			}
			__syncwarp();
			// ---- correlation (sse_conv_cmplx_8n order, convolve_sse_3.c:462-537): 3 outputs per item ----
			const int soff = pk0.y & 0xffff;
			const float2 *hh = c_tab.seq + soff;
			float2 *crow = p.corr + (size_t)b * lmax;
			if (pm1_s && (pk0.y >> 30)) {
				if (hlen == 40) corr_long_items_pm1<40>(dec, pm1_s + soff, len, hlen, lane, crow, NZ);
				else if (hlen == 16) corr_long_items_pm1<16>(dec, pm1_s + soff, len, hlen, lane, crow, NZ);
				else if (hlen == 64) corr_long_items_pm1<64>(dec, pm1_s + soff, len, hlen, lane, crow, NZ);
				else corr_long_items_pm1<0>(dec, pm1_s + soff, len, hlen, lane, crow, NZ);
			} else if (hlen == 40) corr_long_items<40>(dec, hh, len, hlen, lane, crow, NZ);
			else if (hlen == 16) corr_long_items<16>(dec, hh, len, hlen, lane, crow, NZ);
			else if (hlen == 64) corr_long_items<64>(dec, hh, len, hlen, lane, crow, NZ);
			else corr_long_items<0>(dec, hh, len, hlen, lane, crow, NZ);
			__syncwarp();
		}
		pk0 = pk1;
	}
}

// ---------------------------------------------------------------------------------------------
// peak_kernel
// ---------------------------------------------------------------------------------------------
constexpr int kPadRows = 9;    // zero rows on either side of a correlation vector (interpolation reach)
constexpr int kRowPitch = 32;  // float2 per row = one per lane: a lane's accesses stay on its own bank pair whatever the row
// interpolation weights on the 1/512 TOA grid, tap-major with bit-reversed columns: wtab[d * 512 + brev9(F)] =
// sinc(pi * |d - 10 - F/512|).  Bisection step k only ever visits F = odd multiples of 512 / 2^k, which in natural
// column order all fall on ONE shared-memory bank (a 16-way conflict at k = 5); bit reversal spreads every
// step's candidates over consecutive banks, and lanes on the same F still broadcast.
constexpr int kSinc512 = 21 * 512;
__device__ __forceinline__ int brev9(int F) { return (int)(__brev((unsigned)F) >> 23); }

__host__ __device__ inline size_t peak_warp_bytes(int lmax) { return (size_t)(lmax + 2 * kPadRows) * kRowPitch * sizeof(float2); }
__host__ __device__ inline size_t peak_hdr_bytes()
{
	return (size_t)kSinc512 * sizeof(float) + ((SEQ_COUNT * sizeof(SeqInfo) + 15) & ~(size_t)15);
}

// Everything peak detection does for one burst once its correlation vector sits in shared memory (Cl = row 0 of the
// burst's column, rows kRowPitch apart between zero rows): the round bookkeeping, the gates, the TOA bisection, C/I,
// amp.  pwr_at(j) = |decimated sample j|^2 of the burst's correlator window.  Shared by peak_kernel and nb_fused_kernel.
struct PeakRes {
	int rc;
	float2 amp;
	float toa, ci;
	int tsc_out;
	unsigned flags;
	bool write_all; // round 0 defines every output of a valid burst; later rounds only on a hit / error
};
template <class PwrAt>
__device__ __forceinline__ PeakRes peak_lane(const PeakParams &p, const SeqInfo *sinfo, const float *stab, float2 *Cl, bool valid,
					      bool run, const Attempt &at, int type, int tsc, int T, int rc, float2 NZ, PwrAt pwr_at)
{
	unsigned flags = 0;
	float2 amp = make_float2(0.0f, 0.0f);
	float toa = 0.0f, ci = 0.0f;
	int tsc_out = 0;
	bool write_all = false; // round 0 defines every output of a valid burst; later rounds only on a hit / error

	if (valid && p.round == 0) {
		write_all = true;
		if ((type == 1 || type == 5) && tsc > 7) { rc = -3; tsc_out = 0; } // -SIGERR_UNSUPPORTED
		else {
			if (type == 1 || type == 5) tsc_out = tsc;
			if (type_known(type) && T > p.max_toa_bound) rc = -1; // -SIGERR_BOUNDS: caller's bound was wrong
		}
	}
	if (valid && !run && rc == 0 && type_known(type) && !((type == 1 || type == 5) && tsc > 7) && T <= p.max_toa_bound) {
		// the attempt exists but its window is larger than trxb200_detect_config() promised: -SIGERR_BOUNDS
		if (at.seq >= 0 && sinfo[at.seq].len + at.len - 1 > p.ndmax) { rc = -1; write_all = true; }
	}

	if (run) {
		const int len = at.len;
		const SeqInfo si = sinfo[at.seq];
		// rows the interpolation may touch beyond the vector are zero; the vector's last sample is
		// excluded from interpolation (end = size - 1, :1105) and is zeroed once the gates are done
		for (int r = len; r < len + kPadRows && r < p.lmax + kPadRows; r++) Cl[r * kRowPitch] = make_float2(0.0f, 0.0f);
		// fastPeakDetect
		float mx = 0.0f;
		int idx = -1;
		float2 pk = make_float2(0.0f, 0.0f);
		for (int i = 0; i < len; i++) {
			const float2 v = Cl[i * kRowPitch];
			const float pwv = norm2(v);
			if (pwv > mx) { mx = pwv; idx = i; pk = v; }
		}
		float t = (float)idx;
		bool hit = !((t < 3.0f) || (t > (float)(len - 3)));
		if (hit) {
			// computePeakRatio (sps = 1)
			int num = 0;
			float avg = 0.0f;
			for (int i = 2; i <= 5; i++) {
				if (idx - i >= 0) { avg = fa(avg, norm2(Cl[(idx - i) * kRowPitch])); num++; }
				if (idx + i < len) { avg = fa(avg, norm2(Cl[(idx + i) * kRowPitch])); num++; }
			}
			float ratio = 0.0f;
			if (num >= 5) {
				const float rms = (float)((double)sqrtf(avg / (float)num) + 0.00001);
				ratio = sqrtf(norm2(pk)) / rms;
			}
			if (fabsf(ratio - p.thresh) < 1e-5f) flags |= 1u;
			if (ratio < p.thresh) hit = false;
		}
		if (hit) {
			Cl[(len - 1) * kRowPitch] = make_float2(0.0f, 0.0f);
			// peakDetect: early/late bisection; the late point is always early + 2 (:1172), i.e. both sit on
			// the same 1/512 grid position and share their 21 weights.  The bisection starts at t - 1 and moves by
			// 1/2, 1/4, ...: after its first step floor(early) is t - 2 or t - 1 for good, so every interpolation of
			// this burst reads rows t - 12 .. t + 11 only.  They are loaded ONCE into registers (the kernel is bound
			// by shared-memory wavefronts: 23 row loads per step were 60 % of them) and shifted by one row, once,
			// for the lanes whose floor is t - 1.
			float early = t - 1.0f, incr = 0.5f;
			float2 V[24]; // V[k] = row floor(early) - 10 + k once the first step has fixed the floor
			{
				const float2 *cp = Cl + (idx - 12) * kRowPitch;
#pragma unroll
				for (int k = 0; k < 24; k++) V[k] = cp[k * kRowPitch];
			}
			bool first = true, tie_exit = false;
#pragma unroll 1
			for (int it = 0; it < 9; it++) {
				const int m = (int)floorf(early);
				const int F = (int)((early - (float)m) * 512.0f);
				const float *wF = stab + brev9(F);
				float2 e = make_float2(0.0f, 0.0f), l = make_float2(0.0f, 0.0f);
				if (first) {
					// floor = t - 1: rows t - 11 + d are V[d + 1]
#pragma unroll
					for (int d = 0; d < 21; d++) {
						const float w = wF[512 * d];
						e = add2(e, mul2(V[d + 1], bc2(w), NZ));
						l = add2(l, mul2(V[d + 3], bc2(w), NZ));
					}
				} else {
#pragma unroll
					for (int d = 0; d < 21; d++) {
						const float w = wF[512 * d];
						e = add2(e, mul2(V[d], bc2(w), NZ));
						l = add2(l, mul2(V[d + 2], bc2(w), NZ));
					}
				}
				const float ne = norm2(e), nl = norm2(l);
				if (near_tie(ne, nl)) flags |= 2u;
				if (ne < nl) early += incr;
				else if (ne > nl) early -= incr;
				else { tie_exit = true; }
				if (first) {
					first = false;
					// from here on floor(early) is t - 1 (moved up, or stopped on a tie) or t - 2 (moved down)
					if (!(ne > nl)) {
#pragma unroll
						for (int k = 0; k < 23; k++) V[k] = V[k + 1];
					}
				}
				if (tie_exit) break;
				incr *= 0.5f;
			}
			t = early + 1.0f;
			float2 xc = make_float2(0.0f, 0.0f);
			{
				// floor(t) = floor(early) + 1: rows floor(early) - 9 + d are V[d + 1]
				const int m = (int)floorf(t);
				const int F = (int)((t - (float)m) * 512.0f);
				const float *wF = stab + brev9(F);
#pragma unroll
				for (int d = 0; d < 21; d++) {
					const float w = wF[512 * d];
					xc = add2(xc, mul2(V[d + 1], bc2(w), NZ));
				}
			}
			// computeCI
			const int N = si.len;
			const int rt = (int)roundf(t);
			const int ps = at.start + 1 - N + rt;
			if (ps < 0 || ps + N > p.dec_size || rt < 0 || rt >= len) {
				ci = 0.0f;
			} else {
				// S = mean |burst[ps..ps+N)|^2, sequential (:1622-1626); pwr index j = dec index - d0 = rt + k
									float S = 0.0f;
				for (int k0 = 0; k0 < N; k0 += 16) { // N is 16, 40 or 64; sixteen loads in flight, then the ordered sum
					float v[16];
#pragma unroll
					for (int k = 0; k < 16; k++) v[k] = (k0 + k < N) ? pwr_at(rt + k0 + k) : 0.0f;
#pragma unroll
					for (int k = 0; k < 16; k++)
						if (k0 + k < N) S = fa(S, v[k]);
				}
				S = S / (float)N;
				const float Cn = norm2(xc) / si.ci_den;
				ci = fm(3.0103f, log2f(Cn / fs(S, Cn)));
			}
			amp = cmul_exact(xc, make_float2(si.inv_gr, si.inv_gi));
			toa = fs(fs(t, si.toa), (float)at.head);
			rc = at.rc_hit;
			tsc_out = (type == 2 || type == 3) ? p.round : ((type == 1 || type == 5) ? tsc : 0);
			write_all = true;
		}
	}
	if (valid && p.last_round && rc == 0 && type_known(type)) {
		// a further attempt exists but no round was scheduled for it (trxb200_detect_config): never silently skipped
		Attempt nx = make_attempt(type, tsc, T, p.round + 1);
		if (nx.seq >= 0 && !((type == 1 || type == 5) && tsc > 7) && T <= p.max_toa_bound) { rc = -1; write_all = true; }
	}
	PeakRes res;
	res.rc = rc; res.amp = amp; res.toa = toa; res.ci = ci; res.tsc_out = tsc_out; res.flags = flags; res.write_all = write_all;
	return res;
}

__global__ void __launch_bounds__(512, 1)
peak_kernel(PeakParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	float *stab = reinterpret_cast<float *>(smem_raw);
	SeqInfo *sinfo = reinterpret_cast<SeqInfo *>(smem_raw + (size_t)kSinc512 * sizeof(float));
	float2 *C = reinterpret_cast<float2 *>(smem_raw + peak_hdr_bytes() + peak_warp_bytes(p.lmax) * warp);
	const int nrows = p.lmax + 2 * kPadRows;
	const float2 NZ = bc2(p.negzero);

	for (int k = threadIdx.x; k < kSinc512; k += blockDim.x) stab[k] = p.sinc512[k];
	for (int k = threadIdx.x; k < SEQ_COUNT; k += blockDim.x) sinfo[k] = c_tab.info[k];
	for (int k = lane; k < nrows * kRowPitch; k += 32) C[k] = make_float2(0.0f, 0.0f);
	__syncthreads();

	const int ntiles = (p.n + 31) >> 5;
	for (int tile = blockIdx.x * wpb + warp; tile < ntiles; tile += gridDim.x * wpb) {
		const int b = tile * 32 + lane;
		const bool valid = b < p.n;
		int type = 0, tsc = 0, T = 0, rc = 0;
		if (valid) {
			type = load_type(p.type, b, p.sch);
			if (!p.sch) { tsc = p.tsc[b]; T = p.max_toa[b]; }
			if (p.round > 0) rc = p.rc[b];
		}
		Attempt at;
		const bool run = valid && attempt_runs(type, tsc, T, p.max_toa_bound, p.ndmax, p.round, rc, sinfo, at);

		// the warp's NEXT tile is asked for in L2 now (its vectors and powers come from DRAM - the correlator wrote them a whole
		// batch ago - and every lane walks a row of its own: 57 % of the stall samples waited on these loads)
		{
			const long bn = (long)(tile + gridDim.x * wpb) * 32 + lane;
			if (bn < p.n) {
				const char *c0 = reinterpret_cast<const char *>(p.corr + (size_t)bn * p.lmax);
				for (int o = 0; o < p.lmax * 8; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c0 + o));
				const char *w0 = reinterpret_cast<const char *>(p.pwr + (size_t)bn * p.ndmax);
				for (int o = 0; o < p.ndmax * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(w0 + o));
			}
		}
		// ---- bring the tile's correlation vectors in: global [burst][lmax] -> shared [kPadRows + i][lane];
		//      each lane copies its own burst's row (16-byte loads at row stride, conflict-free 8-byte stores) ----
		__syncwarp();
		if (run) {
			const float2 *src = p.corr + (size_t)b * p.lmax;
			float2 *dst = C + kPadRows * kRowPitch + lane;
			const int len = at.len;
			if ((p.lmax & 1) == 0) {
				// twenty samples per round: the ten 16-byte loads are in flight together - a normal burst's whole
				// vector in one round trip (rows are lmax long, so a pair that starts inside the row ends inside it)
				for (int i0 = 0; i0 < len; i0 += 20) {
					float4 v[10];
					if (len > 20 && i0 + 40 <= ((len + 1) & ~1)) {
						// long vectors (access bursts): forty samples per round trip
						float4 u[20];
#pragma unroll
						for (int k = 0; k < 20; k++) u[k] = __ldg(reinterpret_cast<const float4 *>(src + i0 + 2 * k));
#pragma unroll
						for (int k = 0; k < 20; k++) {
							dst[(i0 + 2 * k) * kRowPitch] = make_float2(u[k].x, u[k].y);
							if (i0 + 2 * k + 1 < len) dst[(i0 + 2 * k + 1) * kRowPitch] = make_float2(u[k].z, u[k].w);
						}
						i0 += 20;
						continue;
					}
#pragma unroll
					for (int k = 0; k < 10; k++) {
						v[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
						if (i0 + 2 * k < len) v[k] = __ldg(reinterpret_cast<const float4 *>(src + i0 + 2 * k));
					}
#pragma unroll
					for (int k = 0; k < 10; k++) {
						if (i0 + 2 * k < len) dst[(i0 + 2 * k) * kRowPitch] = make_float2(v[k].x, v[k].y);
						if (i0 + 2 * k + 1 < len) dst[(i0 + 2 * k + 1) * kRowPitch] = make_float2(v[k].z, v[k].w);
					}
				}
			} else {
				for (int i = 0; i < len; i++) dst[i * kRowPitch] = __ldg(&src[i]);
			}
		}
		__syncwarp();

		const float *pwrow = p.pwr + (size_t)b * p.ndmax;
		float2 *Cl0 = C + kPadRows * kRowPitch + lane; // row i of this lane's burst: Cl0[i * kRowPitch]
		const PeakRes res = peak_lane(p, sinfo, stab, Cl0, valid, run, at, type, tsc, T, rc, NZ,
					      [&](int j) { return __ldg(&pwrow[j]); });
		rc = res.rc;
		const float2 amp = res.amp;
		const float toa = res.toa, ci = res.ci;
		const int tsc_out = res.tsc_out;
		const unsigned flags = res.flags;
		const bool write_all = res.write_all;

		if (valid) {
			if (write_all) {
				p.rc[b] = rc;
				reinterpret_cast<float2 *>(p.amp)[b] = amp;
				p.toa[b] = toa;
				p.ci[b] = ci;
				if (p.tsc_out) p.tsc_out[b] = (uint8_t)tsc_out;
			}
			if (p.flags) {
				if (p.round == 0) p.flags[b] = (uint8_t)flags;
				else if (flags) p.flags[b] |= (uint8_t)flags;
			}
		}
	}
}

// maxAmplitude (sigProcLib.cpp:1711-1722) for the bursts detection left at rc == 0: -SIGERR_CLIP is only
// reported when nothing was detected (:1764), and only for the types detectAnyBurst hands to detectGeneralBurst:
// OFF, SCH and unknown types return 0 without a clipping check (:1949-1956).  type == nullptr: every burst is checked.
// One warp per burst.
__global__ void __launch_bounds__(256)
clip_kernel(const float *bursts, int stride, int n, int32_t *rc, uint8_t *flags, const uint8_t *type, int len = 625)
{
	const int lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
		if (rc[b] != 0) continue;
		if (type && !type_known(load_type(type, b, 0))) continue;
		const float2 *x = reinterpret_cast<const float2 *>(bursts) + (size_t)b * stride;
		float mx = 0.0f;
		for (int i = lane; i < len; i += 32) {
			const float2 v = __ldg(&x[i]);
			mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
		}
#pragma unroll
		for (int o = 16; o; o >>= 1)
			mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
		if (lane == 0 && mx > kClipThresh) {
			rc[b] = -2;
			if (flags) flags[b] |= 4;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// detectSCHBurst in its SCH_DETECT_BUFFER state (sigProcLib.cpp:1805-1861): the first acquisition searches a whole
// 12-frame capture.  The capture is decimated 4:1 (15,000 samples), correlated with the 64-symbol sequence at every
// decimated position from start 0 - 960 k complex taps - and the peak logic of detectBurst runs over the 15,000-long
// vector.  Too long for the per-warp shared-memory tiles above, and rare (once per cell acquisition), so it runs from
// global memory (L2 resident: 120 KB per capture) with one thread per output:
//   sch_buf_decim_kernel   out j = sse_conv_real16 order over samples 4j-15 .. 4j (16 zero history samples), stored behind
//                          the 63 zeros convolve() prepends for a CUSTOM span starting at 0 (:312-334)
//   sch_buf_corr_kernel    out i = sse_conv_cmplx_8n order over decimated samples i-63 .. i, stored as column (capture & 31)
//                          of a [kPadRows + len + kPadRows][32] tile: exactly the layout peak_lane() walks
//   sch_buf_peak_kernel    warp = tile, lanes = captures: peak_lane() (argmax, gates, ratio, bisection, C/I, amp)
// ---------------------------------------------------------------------------------------------
struct SchBufDetParams {
	const float *bursts;
	int stride, in_len, len, n;
	float thresh;
	float2 *dec;   // [n][len + 72]: 63 zeros, len samples, 9 zeros
	float2 *ctile; // [tiles][len + 2 * kPadRows][32]
	const float *sinc512;
	int32_t *rc;
	float *amp, *toa, *ci;
	uint8_t *flags;
	float negzero;
};
constexpr int kSchBufDecPad = 72;

__global__ void __launch_bounds__(256)
sch_buf_decim_kernel(SchBufDetParams p)
{
	const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
	if (j >= p.len) return;
	const float2 NZ = bc2(p.negzero);
	const float2 *x = reinterpret_cast<const float2 *>(p.bursts) + (size_t)b * p.stride;
	float2 L[4];
#pragma unroll
	for (int q = 0; q < 4; q++) {
		float2 pr[4];
#pragma unroll
		for (int m = 0; m < 4; m++) {
			const int k = 4 * m + q, idx = 4 * j - 15 + k;
			const float2 v = idx >= 0 ? __ldg(&x[idx]) : make_float2(0.0f, 0.0f);
			pr[m] = mul2(v, bc2(c_tab.dnsamp[k]), NZ);
		}
		L[q] = add2(add2(pr[0], pr[1]), add2(pr[2], pr[3]));
	}
	p.dec[(size_t)b * (p.len + kSchBufDecPad) + 63 + j] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
}

__global__ void __launch_bounds__(256)
sch_buf_corr_kernel(SchBufDetParams p)
{
	const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
	if (i >= p.len) return;
	const float2 NZ = bc2(p.negzero);
	const float2 *dx = p.dec + (size_t)b * (p.len + kSchBufDecPad) + i; // decimated sample i - 63 + k sits at dx[k]
	const float2 *hh = c_tab.seq + c_tab.info[SEQ_SCH].off;
	float2 A[4], B[4];
#pragma unroll
	for (int q = 0; q < 4; q++) { A[q] = make_float2(0.0f, 0.0f); B[q] = make_float2(0.0f, 0.0f); }
#pragma unroll 2
	for (int t0 = 0; t0 < 64; t0 += 8) {
		float2 xw[8];
#pragma unroll
		for (int k = 0; k < 8; k++) xw[k] = dx[t0 + k];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const float2 h1 = hh[t0 + q], h2 = hh[t0 + 4 + q];
			A[q] = add2(A[q], cmul_tap(xw[q], bc2(h1.x), make_float2(h1.y, -h1.y), NZ));
			B[q] = add2(B[q], cmul_tap(xw[4 + q], bc2(h2.x), make_float2(h2.y, -h2.y), NZ));
		}
	}
	float2 L[4];
#pragma unroll
	for (int q = 0; q < 4; q++) L[q] = add2(A[q], B[q]);
	float2 *tile = p.ctile + (size_t)(b >> 5) * (p.len + 2 * kPadRows) * kRowPitch;
	tile[(size_t)(kPadRows + i) * kRowPitch + (b & 31)] = add2(add2(L[0], L[1]), add2(L[2], L[3]));
}

__global__ void __launch_bounds__(32)
sch_buf_peak_kernel(SchBufDetParams q)
{
	const int lane = threadIdx.x, b = blockIdx.x * 32 + lane;
	const bool valid = b < q.n;
	PeakParams p;
	p.n = q.n; p.type = nullptr; p.tsc = nullptr; p.max_toa = nullptr; p.round = 0; p.last_round = 1; p.max_toa_bound = 0;
	p.thresh = q.thresh; p.lmax = q.len; p.ndmax = 1 << 30; p.corr = nullptr; p.pwr = nullptr; p.sinc512 = q.sinc512; p.rc = q.rc;
	p.amp = q.amp; p.toa = q.toa; p.ci = q.ci; p.tsc_out = nullptr; p.flags = q.flags; p.negzero = q.negzero; p.sch = 1;
	p.dec_size = q.len;
	Attempt at;
	at.seq = SEQ_SCH; at.head = 3 + 39 + 64; at.start = 0; at.len = q.len; at.rc_hit = 1; // toa - (3 + 39 + 64), :1853-1854
	float2 *Cl = q.ctile + ((size_t)blockIdx.x * (q.len + 2 * kPadRows) + kPadRows) * kRowPitch + lane;
	const float2 *dec = q.dec + (size_t)(valid ? b : 0) * (q.len + kSchBufDecPad);
	const PeakRes res = peak_lane(p, c_tab.info, q.sinc512, Cl, valid, valid, at, kTypeSchFull, 0, 0, 0, bc2(q.negzero),
				      [&](int j) { return norm2(dec[j]); });
	if (valid) {
		q.rc[b] = res.rc < 0 ? -1 : res.rc;
		reinterpret_cast<float2 *>(q.amp)[b] = res.rc > 0 ? res.amp : make_float2(0.0f, 0.0f);
		q.toa[b] = res.rc > 0 ? res.toa : 0.0f;
		q.ci[b] = res.ci;
		if (q.flags) q.flags[b] = (uint8_t)res.flags;
	}
}

} // namespace trxb200
