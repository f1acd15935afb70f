// Error reporting + ABI version for the pokerrl_b200 C ABI (include/pokerrl_b200.h).
#include <stdio.h>
#include <string.h>

#include "pokerrl_b200.h"
#include "prl_common.cuh"

namespace {
thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
}

namespace prl {
int fail(const char* msg) {
    strncpy(g_err, msg, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
    return -1;
}
int check(cudaError_t e, const char* where) {
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
void count_launch() { ++g_launches; }
}  // namespace prl

extern "C" int prl_abi_version(void) { return PRL_ABI_VERSION; }
extern "C" const char* prl_last_error(void) { return g_err; }
extern "C" unsigned long long prl_launch_count(void) { return g_launches; }
