// misc.cu -- library bookkeeping.
#include "common.cuh"

long long g_b200sp_launches = 0;

extern "C" int b200sp_version(void) { return 1; }
extern "C" int64_t b200sp_launch_count(void) { return (int64_t)g_b200sp_launches; }
