/* shim for <osmocom/core/logging.h>: level constants; LOGP prints to stderr (only reached on write() errors). */
#pragma once
#include <stdio.h>
#define LOGL_DEBUG 1
#define LOGL_INFO 3
#define LOGL_NOTICE 5
#define LOGL_ERROR 7
#define LOGL_FATAL 8
struct log_info;
#define LOGP(cat, level, fmt, args...) fprintf(stderr, fmt, ##args)
