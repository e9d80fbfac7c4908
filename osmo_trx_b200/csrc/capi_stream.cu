// capi_stream.cu — vitac / resampler / filterbank entry points (included by capi.cu)
extern "C" {
int trxb200_vitac_batch(trxb200_ctx *ctx, const float *, int, int, int, int, const uint8_t *, int, int, int, int8_t *, int32_t *, float *, float *) { return fail(ctx, TRXB200_EINVAL, "vitac: not built yet"); }
int trxb200_resampler_create(trxb200_ctx *ctx, int, int, int, float, trxb200_resampler **) { return fail(ctx, TRXB200_EINVAL, "resampler: not built yet"); }
void trxb200_resampler_destroy(trxb200_resampler *) {}
int trxb200_resampler_rotate(trxb200_resampler *, const float *, int, int, float *, int, int, int) { return TRXB200_EINVAL; }
int trxb200_resampler_taps(trxb200_resampler *, int, float *) { return TRXB200_EINVAL; }
int trxb200_channelizer_create(trxb200_ctx *ctx, int, int, int, trxb200_filterbank **) { return fail(ctx, TRXB200_EINVAL, "channelizer: not built yet"); }
int trxb200_synthesis_create(trxb200_ctx *ctx, int, int, int, trxb200_filterbank **) { return fail(ctx, TRXB200_EINVAL, "synthesis: not built yet"); }
void trxb200_filterbank_destroy(trxb200_filterbank *) {}
int trxb200_filterbank_reset(trxb200_filterbank *) { return TRXB200_EINVAL; }
int trxb200_channelizer_rotate(trxb200_filterbank *, const float *, float *, int) { return TRXB200_EINVAL; }
int trxb200_synthesis_rotate(trxb200_filterbank *, const float *, float *, int) { return TRXB200_EINVAL; }
int trxb200_filterbank_taps(trxb200_filterbank *, int, float *) { return TRXB200_EINVAL; }
}
