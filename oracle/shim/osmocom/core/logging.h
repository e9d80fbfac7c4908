/* shim for <osmocom/core/logging.h>: level constants only. */
#pragma once
#define LOGL_DEBUG 1
#define LOGL_INFO 3
#define LOGL_NOTICE 5
#define LOGL_ERROR 7
#define LOGL_FATAL 8
