// oracle/shim/logs/logs.h -- TEST INFRASTRUCTURE. kiwi::logs stand-in: errors go to stderr verbatim, the rest is dropped.
#ifndef ORACLE_SHIM_LOGS_H
#define ORACLE_SHIM_LOGS_H
#include <cstdio>
#ifndef MASTER_PROCESSOR
#define MASTER_PROCESSOR 0
#endif
namespace kiwi {
    namespace logs {
        template<typename... A> inline void v(A...) {}
        template<typename... A> inline void d(A...) {}
        template<typename... A> inline void i(A...) {}
        template<typename... A> inline void s(A...) {}
        template<typename... A> inline void w(A...) {}
        template<typename... A> inline void e(const char *tag, const char *fmt, A...) { fprintf(stderr, "[%s] %s", tag, fmt); }
    }
}
#endif
