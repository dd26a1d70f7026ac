// oracle/shim/spdlog/spdlog.h -- TEST INFRASTRUCTURE: the logging macros the reference uses, as no-ops.
#pragma once
#define SPDLOG_TRACE(...) ((void)0)
#define SPDLOG_DEBUG(...) ((void)0)
#define SPDLOG_INFO(...) ((void)0)
#define SPDLOG_WARN(...) ((void)0)
#define SPDLOG_ERROR(...) ((void)0)
#define SPDLOG_CRITICAL(...) ((void)0)
namespace spdlog {
namespace level { enum level_enum { trace, debug, info, warn, err, critical, off }; }
inline void set_level(level::level_enum) {}
template <class... A> inline void info(A&&...) {}
template <class... A> inline void debug(A&&...) {}
template <class... A> inline void warn(A&&...) {}
template <class... A> inline void error(A&&...) {}
}  // namespace spdlog
