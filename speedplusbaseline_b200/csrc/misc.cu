// misc.cu -- library bookkeeping.
#include "common.cuh"

long long g_b200sp_launches = 0;

extern "C" int b200sp_version(void) { return 1; }
extern "C" int64_t b200sp_launch_count(void) { return (int64_t)g_b200sp_launches; }

bool b200sp_pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
