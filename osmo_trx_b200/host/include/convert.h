/* convert.h — the int16 <-> float entry points of arch/common/convert.h, served by the GPU library.
 * convert_float_short: out = (short)(in * scale) as an SSE3 x86 host computes it (arch/x86/convert.c:63-71): round to
 * nearest even with saturation for whole groups of eight values, the scalar truncating loop for the len % 8 tail;
 * base_convert_float_short: the scalar loop for everything (arch/common/convert_base.c:20-25); convert_short_float /
 * base_convert_short_float: exact. */
#ifndef _CONVERT_H_
#define _CONVERT_H_
#ifdef __cplusplus
extern "C" {
#endif
void convert_float_short(short *out, const float *in, float scale, int len);
void convert_short_float(float *out, const short *in, int len);
void base_convert_float_short(short *out, const float *in, float scale, int len);
void base_convert_short_float(float *out, const short *in, int len);
void convert_init(void);
#ifdef __cplusplus
}
#endif
#endif /* _CONVERT_H_ */
