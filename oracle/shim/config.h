/* oracle/shim/config.h — stand-in for the autoconf-generated config.h of the
 * reference build (configure.ac:91-101,291 cannot run here: no autotools,
 * no libosmocore, no fftw3f).  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_CONFIG_H
#define ORACLE_SHIM_CONFIG_H
#define HAVE_SSE3 1
#define HAVE_SSE4_1 1
#define HAVE___BUILTIN_CPU_SUPPORTS 1
#define PACKAGE_VERSION "1.8.0-oracle"
#endif
