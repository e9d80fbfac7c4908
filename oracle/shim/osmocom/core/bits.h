/* shim for <osmocom/core/bits.h>: the typedefs grgsm_vitac needs and the big-endian stores proto_trxd.c
 * uses (libosmocore declares them in bit16gen.h / bit32gen.h, pulled in by bits.h). */
#pragma once
#include <stdint.h>
typedef int8_t sbit_t;
typedef uint8_t ubit_t;
typedef uint8_t pbit_t;
static inline void osmo_store16be(uint16_t x, void *p)
{
	((uint8_t *)p)[0] = (uint8_t)(x >> 8);
	((uint8_t *)p)[1] = (uint8_t)x;
}
static inline void osmo_store32be(uint32_t x, void *p)
{
	((uint8_t *)p)[0] = (uint8_t)(x >> 24);
	((uint8_t *)p)[1] = (uint8_t)(x >> 16);
	((uint8_t *)p)[2] = (uint8_t)(x >> 8);
	((uint8_t *)p)[3] = (uint8_t)x;
}
