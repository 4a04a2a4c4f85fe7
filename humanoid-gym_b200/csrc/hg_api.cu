// Library-wide symbols of libhg_b200.so.
#include "hg_common.cuh"

thread_local char g_hg_err[512] = "";
std::atomic<int64_t> g_hg_launches{0};

extern "C" int32_t hg_version(void) { return HG_VERSION; }
extern "C" const char* hg_last_error(void) { return g_hg_err; }
extern "C" int64_t hg_launch_count(void) { return g_hg_launches.load(); }
extern "C" int64_t hg_struct_size(int32_t which) {
    switch (which) {
        case 0: return sizeof(HgEnvParams);
        case 1: return sizeof(HgEnvBuffers);
        case 2: return sizeof(HgEnvNoise);
        case 3: return sizeof(HgMlpDesc);
        case 4: return sizeof(HgTransition);
        case 5: return sizeof(HgStorage);
        case 6: return sizeof(HgMiniBatch);
        case 7: return sizeof(HgPpoLossArgs);
        case 8: return sizeof(HgGemm);
        case 9: return sizeof(HgSplit);
        case 10: return sizeof(HgGemmSplit);
        case 11: return sizeof(HgMlpFwdOpts);
        case 12: return sizeof(HgTerrain);
        default: return -1;
    }
}
