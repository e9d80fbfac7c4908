/* osmocom/core/bits.h — the three bit-container typedefs of libosmocore that the burst-DSP interfaces mention
 * (grgsm_vitac.h:27-29 includes this header for sbit_t).  A host that has libosmocore installed uses its header; this
 * stand-in keeps the mirror self-contained where it is absent. */
#pragma once
#include <stdint.h>
typedef int8_t sbit_t;	/* soft bit, -127 .. 127 */
typedef uint8_t ubit_t; /* unpacked bit, 0 / 1 */
typedef uint8_t pbit_t; /* packed bits */
