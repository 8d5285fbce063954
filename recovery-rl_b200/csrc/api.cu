// api.cu -- error plumbing shared by all entry points of librrl.so
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";

void rrl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* rrl_last_error(void) { return g_err; }
extern "C" int rrl_version(void) { return RRL_VERSION; }

// programmatic dependent launch of the vector step's kernels (common.cuh): OFF unless RRL_PDL=1 or rrl_set_pdl(1) -- measured
// on B200 (C4, 65,536 env copies, CUDA graph): 0.366 ms per step with it, 0.339 ms without (profiles/r2/pdl_ab.txt)
static int g_pdl = -1;
int rrl_pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("RRL_PDL");
        g_pdl = (e && e[0] == '1') ? 1 : 0;
    }
    return g_pdl;
}
extern "C" int rrl_set_pdl(int enabled) {
    const int was = rrl_pdl_enabled();
    if (enabled >= 0) g_pdl = enabled ? 1 : 0;   // negative: query only
    return was;
}
