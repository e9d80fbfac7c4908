// capi_stream.cu — vitac / resampler / filterbank entry points (included by capi.cu)

struct trxb200_resampler {
	trxb200_ctx *ctx;
	int p, q, L;
	std::vector<float> taps;
	float *d_taps = nullptr;
};

struct trxb200_filterbank {
	trxb200_ctx *ctx;
	int m, block_len, L, synth;
	std::vector<float> taps;
	float *d_taps = nullptr;
	float2 *d_tw = nullptr;
	float *d_hist[2] = { nullptr, nullptr }; // ping-pong: read by the main kernel, written by the tail kernel
	int cur = 0;
};

// one launch of the MLSE: vitac_lane_kernel (lane = burst) unless switched off; nwin = search windows (0 with a caller's estimate)
static int launch_vitac(trxb200_ctx *ctx, VitacParams &p, int nwin)
{
	cudaStream_t st = ctx->stream;
	if (ctx->tune.vitac_lane) {
		const int N = p.is_ab == 1 ? 88 : 148;
		const size_t smem = (size_t)kVlWarps * vl_warp_bytes(p.cir_in ? 0 : nwin, N);
		if (smem <= 200 * 1024) {
			const int ntiles = (p.n + 31) / 32;
			const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (220 * 1024) / smem));
			const int grid = std::max(1, std::min((ntiles + kVlWarps - 1) / kVlWarps, ctx->sm_count * per_sm));
			auto go = [&](auto kern) {
				if (smem > 48 * 1024) {
					cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
					if (e != cudaSuccess) return fail(ctx, TRXB200_ECUDA, "vitac_lane_kernel attribute", e);
				}
				prof_pre(ctx, st);
				kern<<<grid, kVlWarps * 32, smem, st>>>(p);
				prof_post(ctx, st, "vitac_kernel");
				return post_launch(ctx, "vitac_lane_kernel");
			};
			if (p.cir_in || p.is_ab == 0) return go(vitac_lane_kernel<16, true>);
			if (p.is_ab == 1) return go(vitac_lane_kernel<31, false>);
			return go(vitac_lane_kernel<54, false>);
		}
	}
	const int wpb = 4;
	const size_t smem = (size_t)wpb * vitac_warp_floats(p.nwin_max, p.pitch) * sizeof(float);
	if (smem > 200 * 1024) return fail(ctx, TRXB200_EINVAL, "vitac: clamp range too wide for the on-chip window");
	if (smem > 48 * 1024) CK(cudaFuncSetAttribute(vitac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int grid = std::min(((p.n + 1) / 2 + wpb - 1) / wpb, ctx->sm_count * 8);
	if (grid < 1) grid = 1;
	prof_pre(ctx, st);
	vitac_kernel<<<grid, wpb * 32, smem, st>>>(p);
	prof_post(ctx, st, "vitac_kernel");
	return post_launch(ctx, "vitac_kernel");
}

extern "C" {

int trxb200_vitac_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int n, int is_ab, const uint8_t *tsc,
			int max_delay, int clamp_lo, int clamp_hi, int8_t *bits, int32_t *start, float *corr_max, float *cir)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx) return TRXB200_EINVAL;
	if (!bufs || !bits || !start || !corr_max || n < 0 || max_delay < 0 || max_delay > 64 || is_ab < 0 || is_ab > 2 || (!is_ab && !tsc))
		return fail(ctx, TRXB200_EINVAL, "vitac: bad argument");
	const int N = is_ab == 1 ? 88 : 148, center = is_ab == 1 ? 13 : (is_ab == 2 ? 47 : 66);
	const int s0 = is_ab == 2 ? (center - 10) * 4 : (center - 5) * 4 + 1;
	const int s1 = is_ab == 2 ? (center + 30) * 4 : (center + 10 + (is_ab ? max_delay : 0)) * 4;
	// every read must stay inside the row: [offset+clamp_lo, offset+clamp_hi+4N) and the search windows
	const int tlen = is_ab == 1 ? 31 : (is_ab == 2 ? 54 : 16);
	if (offset + clamp_lo < 0 || offset + clamp_hi + 4 * N > stride || offset + s1 - 1 + 4 * (tlen - 1) >= stride || clamp_lo > clamp_hi)
		return fail(ctx, TRXB200_EINVAL, "vitac: row too short for the search window / clamp range");
	if (n == 0) return TRXB200_OK;
	VitacParams p;
	p.bufs = bufs; p.stride = stride; p.offset = offset; p.n = n; p.is_ab = is_ab; p.tsc = tsc; p.max_delay = max_delay;
	p.clamp_lo = clamp_lo; p.clamp_hi = clamp_hi; p.bits = bits; p.start = start; p.corr_max = corr_max; p.cir = cir;
	p.cir_in = nullptr; p.start_in = nullptr;
	p.nwin_max = s1 - s0;
	p.lo = std::min(clamp_lo, s0);
	p.range = std::max(clamp_hi + 4 * N, s1 + 4 * (tlen - 1)) - p.lo;
	p.pitch = vitac_pitch(p.range);
	return launch_vitac(ctx, p, s1 - s0);
}

/* detect_burst_nb / detect_burst_ab (grgsm_vitac.cpp:105-123) with the CALLER's channel estimate and burst start: matched
 * filter + Viterbi only.  start_in is clamped to [clamp_lo, clamp_hi], which also bounds what is read of each row. */
int trxb200_vitac_detect_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int n, int is_ab, const float *cir_in,
			       const int32_t *start_in, int clamp_lo, int clamp_hi, int8_t *bits)
{
	return trxb200_vitac_detect_ss_batch(ctx, bufs, stride, offset, n, is_ab, cir_in, start_in, clamp_lo, clamp_hi, 3, bits);
}

int trxb200_vitac_detect_ss_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int n, int is_ab, const float *cir_in,
				  const int32_t *start_in, int clamp_lo, int clamp_hi, int start_state, int8_t *bits)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx) return TRXB200_EINVAL;
	if (!bufs || !bits || !cir_in || !start_in || n < 0 || is_ab < 0 || is_ab > 1 || clamp_lo > clamp_hi || start_state < 0 ||
	    start_state > 15)
		return fail(ctx, TRXB200_EINVAL, "vitac_detect: bad argument");
	const int N = is_ab ? 88 : 148;
	if (offset + clamp_lo < 0 || offset + clamp_hi + 4 * N > stride)
		return fail(ctx, TRXB200_EINVAL, "vitac_detect: row too short for the clamp range");
	if (n == 0) return TRXB200_OK;
	VitacParams p;
	p.bufs = bufs; p.stride = stride; p.offset = offset; p.n = n; p.is_ab = is_ab; p.tsc = nullptr; p.max_delay = 0;
	p.clamp_lo = clamp_lo; p.clamp_hi = clamp_hi; p.bits = bits; p.start = nullptr; p.corr_max = nullptr; p.cir = nullptr;
	p.cir_in = cir_in; p.start_in = start_in; p.start_state = start_state;
	p.nwin_max = 20; // cb only holds the 20 taps
	p.lo = clamp_lo;
	p.range = clamp_hi + 4 * N - clamp_lo;
	p.pitch = vitac_pitch(p.range);
	return launch_vitac(ctx, p, 0);
}

/* get_sch_buffer_chan_imp_resp (grgsm_vitac.cpp:298-309) over `len` samples of every row + detect_burst_nb at the position
 * found (ms_rx_lower.cpp:168-177): the first SCH acquisition of a cell. */
int trxb200_vitac_sch_buffer_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int len, int n, int8_t *bits,
				   int32_t *start, float *corr_max, float *cir)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx) return TRXB200_EINVAL;
	const int nwin = len - 64 * 8; // search_stop_pos = len - N_SYNC_BITS * 8
	if (!bufs || !start || !corr_max || !cir || n < 0 || offset < 0 || len <= 0 || offset + len > stride || nwin < 20)
		return fail(ctx, TRXB200_EINVAL, "vitac_sch_buffer: bad argument");
	if (stride - offset < 148 * 4) return fail(ctx, TRXB200_EINVAL, "vitac_sch_buffer: row shorter than a burst");
	if (n == 0) return TRXB200_OK;
	cudaStream_t st = ctx->stream;
	float2 *d_corr = nullptr;
	float *d_pw = nullptr;
	int32_t *d_shift = nullptr;
	CK(cudaMallocAsync(&d_corr, (size_t)n * nwin * sizeof(float2), st));
	CK(cudaMallocAsync(&d_pw, (size_t)n * nwin * sizeof(float), st));
	CK(cudaMallocAsync(&d_shift, (size_t)n * sizeof(int32_t), st));
	SchBufParams q;
	q.bufs = bufs; q.stride = stride; q.offset = offset; q.n = n; q.nwin = nwin; q.corr = d_corr; q.pw = d_pw; q.start = start;
	q.shift = d_shift; q.shift_lo = -offset; q.shift_hi = stride - offset - 148 * 4; q.corr_max = corr_max; q.cir = cir;
	sch_buffer_corr_kernel<<<dim3((nwin + 255) / 256, n), 256, 0, st>>>(q);
	int r = post_launch(ctx, "sch_buffer_corr_kernel");
	if (!r) {
		sch_buffer_window_kernel<<<n, 256, 0, st>>>(q);
		r = post_launch(ctx, "sch_buffer_window_kernel");
	}
	if (!r && bits) {
		VitacParams p;
		p.bufs = bufs; p.stride = stride; p.offset = offset; p.n = n; p.is_ab = 0; p.tsc = nullptr; p.max_delay = 0;
		p.clamp_lo = 0; p.clamp_hi = 0; p.bits = bits; p.start = nullptr; p.corr_max = nullptr; p.cir = nullptr;
		p.cir_in = cir; p.start_in = nullptr; p.row_shift = d_shift; p.start_state = 3;
		p.nwin_max = 20;
		p.lo = 0;
		p.range = 4 * 148;
		p.pitch = vitac_pitch(p.range);
		r = launch_vitac(ctx, p, 0);
	}
	cudaFreeAsync(d_corr, st);
	cudaFreeAsync(d_pw, st);
	cudaFreeAsync(d_shift, st);
	return r;
}

/* ---------------- Resampler ---------------- */
int trxb200_resampler_create(trxb200_ctx *ctx, int p, int q, int filt_len, float bw, trxb200_resampler **out)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !out) return TRXB200_EINVAL;
	*out = nullptr;
	if (p <= 0 || q <= 0 || filt_len <= 0 || filt_len > 32) // Resampler::init returns false for zero sizes
		return fail(ctx, TRXB200_EINVAL, "resampler: bad p/q/filt_len");
	trxb200_resampler *r = new trxb200_resampler();
	r->ctx = ctx; r->p = p; r->q = q; r->L = filt_len;
	build_resampler_taps(p, q, filt_len, bw, r->taps);
	cudaError_t e = cudaMalloc(&r->d_taps, r->taps.size() * sizeof(float));
	if (e == cudaSuccess) e = cudaMemcpy(r->d_taps, r->taps.data(), r->taps.size() * sizeof(float), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { delete r; return fail(ctx, TRXB200_ECUDA, "resampler_create", e); }
	*out = r;
	return TRXB200_OK;
}

void trxb200_resampler_destroy(trxb200_resampler *r)
{
	if (!r) return;
	DevGuard dg(r->ctx->device);
	cudaFree(r->d_taps);
	delete r;
}

int trxb200_resampler_taps(trxb200_resampler *r, int path, float *out_host)
{
	if (!r || !out_host || path < 0 || path >= r->p) return TRXB200_EINVAL;
	std::memcpy(out_host, r->taps.data() + (size_t)path * r->L, sizeof(float) * r->L);
	return r->L;
}

static int resampler_rotate_impl(trxb200_resampler *r, const float *in, int in_len, int in_stride, float *out, int out_len,
				 int out_stride, int n_streams, bool capped);
// resampler_pq_kernel (resampler_pq.cu): the ratio is a template parameter
extern "C++" {
template <int P, int Q, int R, int NCH>
static int launch_resampler_pq(trxb200_resampler *r, const float *in, int in_stride, float *out, int out_len, int out_stride, int n_streams)
{
	trxb200_ctx *ctx = r->ctx;
	using G = RsPqGeom<P, Q, R, NCH>;
	static_assert(sizeof(RsPqParams<P, Q>) <= 32764, "kernel parameter block");
	RsPqParams<P, Q> M;
	M.in = in; M.out = out; M.in_stride = in_stride; M.out_len = out_len; M.out_stride = out_stride; M.n_streams = n_streams;
	M.negzero = -0.0f;
	for (int rho = 0; rho < P; rho++)
		std::memcpy(M.tp[rho], r->taps.data() + (size_t)((Q * rho) % P) * 16, 16 * sizeof(float));
	const int warps = (int)std::max<size_t>(1, std::min<size_t>(12, (size_t)(225 * 1024 - G::hdr_bytes) / G::warp_bytes));
	const size_t smem = G::hdr_bytes + (size_t)warps * G::warp_bytes;
	bool &cfg = ctx->cfg_rspq[P > Q ? 0 : 1]; // (the attribute is per device: a flag per context, not per process)
	if (!cfg) {
		CK(cudaFuncSetAttribute(resampler_pq_kernel<P, Q, R, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
		cfg = true;
	}
	const long tiles = (long)n_streams * ((out_len / P + 31) / 32);
	const int grid = (int)std::max<long>(1, std::min<long>((tiles + warps - 1) / warps, (long)ctx->sm_count));
	prof_pre(ctx, ctx->stream);
	resampler_pq_kernel<P, Q, R, NCH><<<grid, warps * 32, smem, ctx->stream>>>(M);
	prof_post(ctx, ctx->stream, "resampler_pq_kernel");
	return post_launch(ctx, "resampler_pq_kernel");
}
} // extern "C++"
int trxb200_resampler_rotate(trxb200_resampler *r, const float *in, int in_len, int in_stride, float *out, int out_len,
			     int out_stride, int n_streams)
{
	return resampler_rotate_impl(r, in, in_len, in_stride, out, out_len, out_stride, n_streams, true);
}
int trxb200_resampler_rotate_stream(trxb200_resampler *r, const float *in, int in_len, int in_stride, float *out, int out_len,
				    int out_stride, int n_streams)
{
	return resampler_rotate_impl(r, in, in_len, in_stride, out, out_len, out_stride, n_streams, false);
}
static int resampler_rotate_impl(trxb200_resampler *r, const float *in, int in_len, int in_stride, float *out, int out_len,
				 int out_stride, int n_streams, bool capped)
{
	if (!r) return TRXB200_EINVAL;
	trxb200_ctx *ctx = r->ctx;
	DevGuard dg(ctx->device);
	if (!in || !out || n_streams < 0 || in_len <= 0 || out_len <= 0)
		return fail(ctx, TRXB200_EINVAL, "resampler_rotate: bad argument");
	// check_vec_len (Resampler.cpp:98-129) + MAX_OUTPUT_LEN; the stream form only keeps the whole-period condition
	if (in_len % r->q || out_len % r->p || in_len / r->q != out_len / r->p || (capped && out_len > 4096 * 4))
		return fail(ctx, TRXB200_EINVAL, "resampler_rotate: block length mismatch");
	if (n_streams == 0) return TRXB200_OK;
	if (r->L == 16 && ctx->tune.resamp_pq) {
		// the multi-ARFCN interface's own ratios: R outputs per thread from one register window, everything compile-time
		if (r->p == 65 && r->q == 48) return launch_resampler_pq<65, 48, 5, 3>(r, in, in_stride, out, out_len, out_stride, n_streams);
		if (r->p == 48 && r->q == 65) return launch_resampler_pq<48, 65, 4, 2>(r, in, in_stride, out, out_len, out_stride, n_streams);
	}
	if (r->L == 16 && r->q <= r->p && r->p <= kRsMaxP && ctx->tune.resamp_up) {
		// interpolating ratio: three outputs per thread from one register window (filterbank.cu, resampler_up_kernel)
		static_assert(sizeof(ResampUpParams) <= 32764, "kernel parameter block");
		ResampUpParams P;
		P.in = in; P.out = out; P.in_stride = in_stride; P.out_len = out_len; P.out_stride = out_stride; P.n_streams = n_streams;
		P.p = r->p; P.q = r->q; P.negzero = -0.0f;
		rs_up_fill(P, r->taps.data());
		const size_t smem = rs_up_smem(r->p, r->q);
		CK(cudaFuncSetAttribute(resampler_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		const long tiles = (long)n_streams * ((out_len / r->p + kRsTile - 1) / kRsTile);
		if (tiles > 0x7fffffffL) return fail(ctx, TRXB200_EINVAL, "resampler_rotate: too many tiles");
		const int grid = (int)std::max<long>(1, std::min<long>(tiles, (long)ctx->sm_count * 2));
		prof_pre(ctx, ctx->stream);
		resampler_up_kernel<<<grid, 256, smem, ctx->stream>>>(P);
		prof_post(ctx, ctx->stream, "resampler_up_kernel");
		return post_launch(ctx, "resampler_up_kernel");
	}
	if (r->L == 16 && r->p <= 256 && r->q < 2 * r->p) {
		// shared-memory staged, taps-in-registers kernel: P polyphase periods of one stream per tile (about 3,000
		// input samples, so that a tile's compute phase is long against the latency of its staging loads)
		const int P = std::max(1, std::min(4096 / r->p, 3000 / r->q));
		const int slots = P * r->q + 15;
		const size_t smem = (size_t)2 * slots * sizeof(float2); // two window buffers
		const long tiles = (long)n_streams * ((out_len / r->p + P - 1) / P);
		const int grid = (int)std::max<long>(1, std::min<long>(tiles, (long)ctx->sm_count * 4));
		prof_pre(ctx, ctx->stream);
		resampler16_kernel<<<grid, 256, smem, ctx->stream>>>(in, in_stride, out, out_len, out_stride, n_streams, r->p, r->q, P,
								     r->d_taps, -0.0f);
		prof_post(ctx, ctx->stream, "resampler16_kernel");
		return post_launch(ctx, "resampler16_kernel");
	}
	resampler_kernel<<<grid_for(ctx, (long)n_streams * out_len, 256, 8), 256, 0, ctx->stream>>>(
		in, in_stride, out, out_len, out_stride, n_streams, r->p, r->q, r->L, r->d_taps);
	return post_launch(ctx, "resampler_kernel");
}

/* ---------------- Channelizer / Synthesis ---------------- */
static int fb_create(trxb200_ctx *ctx, int m, int block_len, int h_len, int synth, trxb200_filterbank **out)
{
	if (!ctx || !out) return TRXB200_EINVAL;
	DevGuard dg(ctx->device);
	*out = nullptr;
	if (m < 1 || m > 256 || h_len < 1 || h_len > 32 || block_len < h_len)
		return fail(ctx, TRXB200_EINVAL, "filterbank: bad m/block_len/h_len");
	trxb200_filterbank *fb = new trxb200_filterbank();
	fb->ctx = ctx; fb->m = m; fb->block_len = block_len; fb->L = h_len; fb->synth = synth;
	build_channelizer_taps(m, h_len, fb->taps);
	std::vector<float2> tw(m);
	for (int k = 0; k < m; k++) {
		const double ph = -2.0 * M_PI * (double)k / (double)m; // forward DFT, both directions (ChannelizerBase.cpp:154)
		tw[k] = make_float2((float)cos(ph), (float)sin(ph));
	}
	cudaError_t e = cudaMalloc(&fb->d_taps, fb->taps.size() * sizeof(float));
	if (e == cudaSuccess) e = cudaMemcpy(fb->d_taps, fb->taps.data(), fb->taps.size() * sizeof(float), cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMalloc(&fb->d_tw, m * sizeof(float2));
	if (e == cudaSuccess) e = cudaMemcpy(fb->d_tw, tw.data(), m * sizeof(float2), cudaMemcpyHostToDevice);
	for (int k = 0; k < 2 && e == cudaSuccess; k++) {
		e = cudaMalloc(&fb->d_hist[k], (size_t)m * h_len * 8);
		if (e == cudaSuccess) e = cudaMemset(fb->d_hist[k], 0, (size_t)m * h_len * 8);
	}
	if (e != cudaSuccess) { trxb200_filterbank_destroy(fb); return fail(ctx, TRXB200_ECUDA, "filterbank_create", e); }
	*out = fb;
	return TRXB200_OK;
}

int trxb200_channelizer_create(trxb200_ctx *ctx, int m, int block_len, int h_len, trxb200_filterbank **out)
{
	return fb_create(ctx, m, block_len, h_len, 0, out);
}
int trxb200_synthesis_create(trxb200_ctx *ctx, int m, int block_len, int h_len, trxb200_filterbank **out)
{
	return fb_create(ctx, m, block_len, h_len, 1, out);
}

void trxb200_filterbank_destroy(trxb200_filterbank *fb)
{
	if (!fb) return;
	DevGuard dg(fb->ctx->device);
	cudaFree(fb->d_taps); cudaFree(fb->d_tw); cudaFree(fb->d_hist[0]); cudaFree(fb->d_hist[1]);
	delete fb;
}

int trxb200_filterbank_reset(trxb200_filterbank *fb)
{
	if (!fb) return TRXB200_EINVAL;
	trxb200_ctx *ctx = fb->ctx;
	DevGuard dg(ctx->device);
	for (int k = 0; k < 2; k++) CK(cudaMemsetAsync(fb->d_hist[k], 0, (size_t)fb->m * fb->L * 8, ctx->stream));
	return TRXB200_OK;
}

int trxb200_filterbank_taps(trxb200_filterbank *fb, int branch, float *out_host)
{
	if (!fb || !out_host || branch < 0 || branch >= fb->m) return TRXB200_EINVAL;
	std::memcpy(out_host, fb->taps.data() + (size_t)branch * fb->L, sizeof(float) * fb->L);
	return fb->L;
}

int trxb200_channelizer_rotate(trxb200_filterbank *fb, const float *in, float *out, int n_blocks)
{
	if (!fb) return TRXB200_EINVAL;
	if (n_blocks < 0) return fail(fb->ctx, TRXB200_EINVAL, "channelizer_rotate: bad argument");
	const long total_t = (long)n_blocks * fb->block_len;
	return trxb200_channelizer_rotate_strided(fb, in, total_t, out, total_t);
}

int trxb200_channelizer_prime(trxb200_filterbank *fb, const float *prev_in, long n_prev_t)
{
	if (!fb) return TRXB200_EINVAL;
	trxb200_ctx *ctx = fb->ctx;
	DevGuard dg(ctx->device);
	if (fb->synth || !prev_in || n_prev_t < fb->L) return fail(ctx, TRXB200_EINVAL, "channelizer_prime: bad argument");
	channelizer_hist_kernel<<<(fb->m * fb->L + 127) / 128, 128, 0, ctx->stream>>>(prev_in, fb->d_hist[fb->cur ^ 1], fb->m, fb->L, n_prev_t);
	fb->cur ^= 1;
	return post_launch(ctx, "channelizer_hist_kernel");
}

int trxb200_channelizer_rotate_strided(trxb200_filterbank *fb, const float *in, long total_t, float *out, long out_stride)
{
	if (!fb) return TRXB200_EINVAL;
	trxb200_ctx *ctx = fb->ctx;
	DevGuard dg(ctx->device);
	if (fb->synth || !in || !out || total_t < 0 || out_stride < total_t || (total_t > 0 && total_t < fb->L))
		return fail(ctx, TRXB200_EINVAL, "channelizer_rotate: bad argument");
	if (total_t == 0) return TRXB200_OK;
	int r;
	if (fb->m == 64 && fb->L == 16 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
		// 8 x 8 split transform + register sliding-window FIRs (filterbank.cu)
		if (!ctx->cfg_ch64) {
			CK(cudaFuncSetAttribute(channelizer64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCh64Smem));
			ctx->cfg_ch64 = true;
		}
		const int grid = (int)std::min<long>((total_t + kFbT - 1) / kFbT, (long)ctx->sm_count * 3);
		prof_pre(ctx, ctx->stream);
		channelizer64_kernel<<<grid, 256, kCh64Smem, ctx->stream>>>(in, fb->d_hist[fb->cur], out, out_stride, total_t, fb->d_taps, fb->d_tw);
		prof_post(ctx, ctx->stream, "channelizer64_kernel");
		r = post_launch(ctx, "channelizer64_kernel");
	} else if (fb->L == 16 && (fb->m == 4 || fb->m == 8 || fb->m == 16)) {
		if (fb->m == 4) launch_channelizer_small<4>(ctx->sm_count, ctx->stream, in, fb->d_hist[fb->cur], out, out_stride, total_t, fb->d_taps, fb->d_tw);
		else if (fb->m == 8) launch_channelizer_small<8>(ctx->sm_count, ctx->stream, in, fb->d_hist[fb->cur], out, out_stride, total_t, fb->d_taps, fb->d_tw);
		else launch_channelizer_small<16>(ctx->sm_count, ctx->stream, in, fb->d_hist[fb->cur], out, out_stride, total_t, fb->d_taps, fb->d_tw);
		r = post_launch(ctx, "channelizer_small_kernel");
	} else {
		const size_t smem = ((size_t)fb->m * 33 + fb->m) * sizeof(float2);
		if (smem > 48 * 1024) CK(cudaFuncSetAttribute(channelizer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		const int grid = (int)std::min<long>((total_t + 31) / 32, (long)ctx->sm_count * 4);
		channelizer_kernel<<<grid, 256, smem, ctx->stream>>>(in, fb->d_hist[fb->cur], out, out_stride, fb->m, fb->L, total_t, fb->d_taps, fb->d_tw);
		r = post_launch(ctx, "channelizer_kernel");
	}
	if (r) return r;
	channelizer_hist_kernel<<<(fb->m * fb->L + 127) / 128, 128, 0, ctx->stream>>>(in, fb->d_hist[fb->cur ^ 1], fb->m, fb->L, total_t);
	fb->cur ^= 1;
	return post_launch(ctx, "channelizer_hist_kernel");
}

int trxb200_synthesis_rotate(trxb200_filterbank *fb, const float *in, float *out, int n_blocks)
{
	if (!fb) return TRXB200_EINVAL;
	trxb200_ctx *ctx = fb->ctx;
	DevGuard dg(ctx->device);
	if (!fb->synth || !in || !out || n_blocks < 0) return fail(ctx, TRXB200_EINVAL, "synthesis_rotate: bad argument");
	if (n_blocks == 0) return TRXB200_OK;
	const long total_t = (long)n_blocks * fb->block_len;
	int r;
	if (fb->m == 64 && fb->L == 16) {
		if (!ctx->cfg_sy64) {
			CK(cudaFuncSetAttribute(synthesis64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSy64Smem));
			ctx->cfg_sy64 = true;
		}
		const long ntiles = (total_t + kFbT - 1) / kFbT;
		const long want = std::min<long>(ntiles, (long)ctx->sm_count * 3);
		const int per = (int)((ntiles + want - 1) / want); // consecutive tiles per CTA (the FIR halo is carried on chip)
		const int grid = (int)((ntiles + per - 1) / per);
		synthesis64_kernel<<<grid, 256, kSy64Smem, ctx->stream>>>(in, fb->d_hist[fb->cur], out, total_t, per, fb->d_taps, fb->d_tw);
		r = post_launch(ctx, "synthesis64_kernel");
	} else if (fb->L == 16 && (fb->m == 4 || fb->m == 8 || fb->m == 16)) {
		if (fb->m == 4) launch_synthesis_small<4>(ctx->sm_count, ctx->stream, in, fb->d_hist[fb->cur], out, total_t, fb->d_taps, fb->d_tw);
		else if (fb->m == 8) launch_synthesis_small<8>(ctx->sm_count, ctx->stream, in, fb->d_hist[fb->cur], out, total_t, fb->d_taps, fb->d_tw);
		else launch_synthesis_small<16>(ctx->sm_count, ctx->stream, in, fb->d_hist[fb->cur], out, total_t, fb->d_taps, fb->d_tw);
		r = post_launch(ctx, "synthesis_small_kernel");
	} else {
		const size_t smem = ((size_t)fb->m * 65 * 2 + fb->m) * sizeof(float2);
		if (smem > 48 * 1024) CK(cudaFuncSetAttribute(synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		const int grid = (int)std::min<long>((total_t + 31) / 32, (long)ctx->sm_count * 2);
		synthesis_kernel<<<grid, 256, smem, ctx->stream>>>(in, fb->d_hist[fb->cur], out, fb->m, fb->L, total_t, fb->d_taps, fb->d_tw);
		r = post_launch(ctx, "synthesis_kernel");
	}
	if (r) return r;
	synthesis_tail_kernel<<<(fb->m * fb->L + 127) / 128, 128, 0, ctx->stream>>>(in, fb->d_hist[fb->cur ^ 1], fb->m, fb->L, total_t);
	fb->cur ^= 1;
	return post_launch(ctx, "synthesis_tail_kernel");
}

} // extern "C"
