/* shim for <osmocom/core/endian.h>: the oracle build targets little-endian hosts (x86-64 / aarch64). */
#pragma once
#define OSMO_IS_LITTLE_ENDIAN 1
#define OSMO_IS_BIG_ENDIAN 0
