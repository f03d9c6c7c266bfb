#include "yb_common.h"

namespace yb {

std::string& last_error_slot() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_slot() = buf;
    return code;
}

}  // namespace yb

extern "C" int yb_abi_version(void) { return YB_ABI_VERSION; }
extern "C" const char* yb_last_error(void) { return yb::last_error_slot().c_str(); }
