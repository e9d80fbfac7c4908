/* oracle/shim/Logger.h — shadows CommonLibs/Logger.h (which needs libosmocore)
 * for the oracle build.  LOG(level) << ... becomes a sink that discards its
 * operands; the hot path only logs on clip / bad type / resampler init
 * (sigProcLib.cpp:1748,1950,2163).  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_LOGGER_H
#define ORACLE_SHIM_LOGGER_H
#include <ostream>
#include <sstream>
struct OracleNullLog {
	std::ostringstream os;
	std::ostream &get() { return os; }
};
#define LOG(level) OracleNullLog().get()
#define LOGC(cat, level) OracleNullLog().get()
#define LOGCHAN(chan, cat, level) OracleNullLog().get()
#endif
