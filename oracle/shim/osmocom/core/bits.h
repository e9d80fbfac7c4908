/* shim for <osmocom/core/bits.h>: only the typedefs grgsm_vitac needs. */
#pragma once
#include <stdint.h>
typedef int8_t sbit_t;
typedef uint8_t ubit_t;
typedef uint8_t pbit_t;
