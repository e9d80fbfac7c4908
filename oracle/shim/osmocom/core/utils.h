/* shim for <osmocom/core/utils.h>: nothing from it is used on the hot path. */
#pragma once
