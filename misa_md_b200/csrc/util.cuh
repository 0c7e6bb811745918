// misa_md_b200/csrc/util.cuh -- error plumbing shared by the host-side translation units.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/misa_b200.h"

// -------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t _e = (call);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return fail(MISA_B200_ENODEV, std::string(#call) + ": " + cudaGetErrorString(_e));               \
    } while (0)
#define REQ(cond, code, msg)                 \
    do {                                     \
        if (!(cond)) return fail(code, msg); \
    } while (0)
#define TRY(call)                \
    do {                         \
        int _rc = (call);        \
        if (_rc != 0) return _rc; \
    } while (0)

