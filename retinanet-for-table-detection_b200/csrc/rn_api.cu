// Library-level entry points: version and per-thread error text.
#include "rn_common.cuh"

static thread_local char g_error[512] = "";

char* rn_error_buffer() { return g_error; }

extern "C" int rn_version(void) { return 100; }   // 0.1.0

extern "C" const char* rn_last_error(void) { return g_error; }
