// Error reporting and ABI version of libgpet_b200.so.
#include <stdarg.h>
#include <string.h>

#include "gpet_common.cuh"

namespace gpet {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gpet

extern "C" const char* gpet_last_error(void) { return gpet::g_err; }
extern "C" int gpet_abi_version(void) { return 1; }
