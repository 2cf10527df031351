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

namespace gpet {
// Tuning knobs (launch shapes / kernel variants); defaults are the measured best on B200.
int g_tune[GPET_TUNE_COUNT] = {128, 1, 0, 0, 4, 4, 1, 0, 64, 0, 32, 512, 2, 1};
}  // namespace gpet

extern "C" int gpet_set_tuning(int knob, int value) {
    if (knob < 0 || knob >= GPET_TUNE_COUNT) {
        gpet::set_error("gpet_set_tuning: unknown knob %d", knob);
        return GPET_ERR_INVALID;
    }
    gpet::g_tune[knob] = value;
    return GPET_OK;
}

extern "C" const char* gpet_last_error(void) { return gpet::g_err; }
extern "C" int gpet_abi_version(void) { return GPET_ABI_VERSION; }
