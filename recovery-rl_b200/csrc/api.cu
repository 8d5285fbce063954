// api.cu -- error plumbing shared by all entry points of librrl.so
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void rrl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* rrl_last_error(void) { return g_err; }
extern "C" int rrl_version(void) { return RRL_VERSION; }
