/* shim for <osmocom/core/panic.h>: osmo_panic -> print + abort. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
static inline void osmo_panic(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
	abort();
}
