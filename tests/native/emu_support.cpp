// TEST INFRASTRUCTURE — the two symbols arah_api.cu provides to the other translation units of libarah_b200.so.
#include <string>
static thread_local std::string g_err;
extern "C" int arah_internal_fail(int code, const char* msg) { g_err = msg ? msg : ""; return code; }
extern "C" const char* arah_last_error(void) { return g_err.c_str(); }
