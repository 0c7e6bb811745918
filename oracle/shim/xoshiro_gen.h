// oracle/shim/xoshiro_gen.h -- TEST INFRASTRUCTURE. Declaration only: the build selects RAND_MT (the reference
// default, config.cmake:22), so the xoshiro generator is never instantiated.
#ifndef ORACLE_SHIM_XOSHIRO_H
#define ORACLE_SHIM_XOSHIRO_H
namespace util { namespace random { class xoroshiro128_plus; } }
#endif
