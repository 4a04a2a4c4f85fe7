// Shared device/host helpers for libhg_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "hg_b200.h"

#define HG_NUM_SMS 148

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
extern thread_local char g_hg_err[512];
extern std::atomic<int64_t> g_hg_launches;

static inline int32_t hg_fail(int32_t code, const char* what) {
    snprintf(g_hg_err, sizeof(g_hg_err), "%s", what);
    return code;
}
static inline int32_t hg_cuda_status(const char* where) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        snprintf(g_hg_err, sizeof(g_hg_err), "%s: %s", where, cudaGetErrorString(e));
        cudaGetLastError();
        return (int32_t)e;
    }
    return 0;
}
#define HG_REQUIRE(ptr)                                                         \
    do {                                                                        \
        if ((ptr) == nullptr) return hg_fail(HG_E_NULL, #ptr " is NULL");       \
    } while (0)
#define HG_LAUNCHED(n) g_hg_launches.fetch_add((n), std::memory_order_relaxed)

static inline bool hg_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: no state in HBM.
// ---------------------------------------------------------------------------
struct HgPhilox {
    uint32_t c[4];
};
__device__ __forceinline__ HgPhilox hg_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    HgPhilox o; o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
    return o;
}
// uniform in [0,1) with 24 random bits (same support as torch.rand for fp32)
__device__ __forceinline__ float hg_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// standard normal from two 32-bit words (Box-Muller; u1 in (0,1])
__device__ __forceinline__ float hg_normal(uint32_t a, uint32_t b) {
    float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = hg_u01(b);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
// stream ids for the env-side draws (c2 of the Philox counter)
enum : uint32_t {
    HG_RNG_CMD_CB = 1, HG_RNG_CMD_RS = 2, HG_RNG_DOF = 3, HG_RNG_PUSH = 4, HG_RNG_OBS = 5,
    HG_RNG_DELAY = 6, HG_RNG_ACT = 7, HG_RNG_SAMPLE = 8, HG_RNG_LEVEL = 9, HG_RNG_ROOT = 10
};
