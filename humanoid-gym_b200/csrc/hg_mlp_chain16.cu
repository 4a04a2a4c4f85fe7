// PPO.act in ONE launch, fp16x3 operands: the same persistent multi-layer tile kernel as hg_mlp_chain.cu (actor AND critic,
// actor_critic.py:54-77,111-128; ppo.py:91-101; layer-to-layer dependencies resolved on the device), but every operand
// travels as two fp16 planes x ~= hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 significant bits, the same as the hi / lo
// pair of 3xTF32) and the product is three tcgen05 kind::f16 MMAs  D += A_lo B_hi + A_hi B_lo + A_hi B_hi  with fp32
// accumulation.  Against 3xTF32 on fp32 tiles that halves the operand bytes per k (the shared-memory / L2 feed is what
// bounds these main loops) and doubles the MMA rate; network-level error is the same ~1e-7
// (tests/test_ppo_gpu.py holds the rollout to the 1e-5 bar against the fp64 reference).
//
// fp16 has a narrow exponent: weights are stored scaled by 2^10 (their lo planes would otherwise be fp16 subnormals:
// measured 50x the error at network level) and the epilogue multiplies the accumulator by 2^-10 (exact); activations are
// O(1) (observations are clipped to +-18 by the env, hidden units are ELU outputs) and saturate at 65504.
//
// Data flow: split_inputs_kernel writes the fp16 planes of obs / critic obs, every hidden layer's epilogue writes its
// activations as planes (shared-memory staging + one TMA store per 32 x 32 block, as hg_gemm_bf3.cu), the output layers
// write fp32 mean / value and the actor's samples the action (same arithmetic as policy_sample_kernel).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "hg_common.cuh"
#include "hg_tc_ptx.cuh"

using namespace hgtc;

namespace {

constexpr int BM = 128, BK = 64, BN_MAX = 128;
constexpr int STAGES = 3;
constexpr int A_PLANE = BM * BK * 2;                             // 16 KB: one fp16 plane of the A tile
constexpr int A_BYTES = 2 * A_PLANE;
constexpr int STAGE_BYTES = A_BYTES + 2 * BN_MAX * BK * 2;       // {A_hi, A_lo, B_hi, B_lo} = 64 KB per 64-k block
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int EPI_STAGE_WARP = 2 * 32 * 64;                      // [2 planes][32 rows][32 fp16], SWIZZLE_64B
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * EPI_STAGE_WARP + 1024 + 256;
constexpr int MAX_CHAIN = 8;
constexpr float kLogSqrt2Pi = 0.9189385332046727f;
constexpr float kWScale = 1024.0f, kWScaleInv = 1.0f / 1024.0f;

enum { CH_BIAS = 1, CH_BIAS_ELU = 2, CH_BIAS_SAMPLE = 5 };

struct Chain16Layer {
    float* C; const float* bias; int64_t ldc;                   // C: fp32 output (last layer of a net); hidden layers leave through maps.c
    int N, K, BN, epi, tiles_n, dep, rot;
};
struct Chain16Args {
    int M, tiles_m, n_layers;
    int* counters;                       // [MAX_CHAIN][tiles_m] tile counters + "CTAs done"
    Chain16Layer L[MAX_CHAIN];
    const float* stdv; const float* eps; float* actions; float* logp; float* sigma;
    uint64_t seed, step; const uint64_t* step_dev;
};
struct alignas(64) Chain16Maps { CUtensorMap a[MAX_CHAIN], b[MAX_CHAIN], c[MAX_CHAIN]; };

// K-major tile [rows][64 x 16 bit], SWIZZLE_128B: SBO = 1024 B (8 rows x 128 B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int N) {                 // kind::f16: D = F32, A = B = F16, K-major, M = 128
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// (x0 at the lower address); satfinite: |x| > 65504 saturates instead of becoming inf
__device__ __forceinline__ uint32_t pack_f16x2(float x0, float x1) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t p) {
    return __half22float2(*reinterpret_cast<const __half2*>(&p));
}

__global__ void __launch_bounds__(THREADS, 1) mlp_chain16_kernel(const __grid_constant__ Chain16Maps maps, const Chain16Args g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* epi_stage = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + EPI_WARPS * EPI_STAGE_WARP);
    uint64_t* full = bars;                          // [S] TMA -> MMA
    uint64_t* empty = full + STAGES;                // [S] MMA -> TMA
    uint64_t* tmem_full = empty + STAGES;           // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;           // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == EPI_WARPS + 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == EPI_WARPS) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int l = 0; l < g.n_layers; ++l) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[l]) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[l]) : "memory");
            }
            int it = 0;
            for (int l = 0; l < g.n_layers; ++l) {
                const Chain16Layer& Ly = g.L[l];
                const int items = g.tiles_m * Ly.tiles_n, num_kb = (Ly.K + BK - 1) / BK;
                const uint32_t tx = (uint32_t)A_BYTES + 2u * (uint32_t)Ly.BN * BK * 2u;
                for (int w = (blockIdx.x + Ly.rot) % G; w < items; w += G) {
                    const int tm = w / Ly.tiles_n, tn = w - tm * Ly.tiles_n;
                    // The WEIGHT tiles of the item's first stages depend on nothing: they are requested before the wait for the
                    // producing layer, so only the activation tiles pay their L2 latency after the dependency resolves.
                    const int pre = (Ly.dep >= 0) ? min(num_kb, STAGES) : 0;
                    for (int kb = 0; kb < pre; ++kb) {
                        const int s = (it + kb) % STAGES;
                        mbar_wait(&empty[s], (((it + kb) / STAGES) & 1) ^ 1);
                        mbar_expect_tx(&full[s], tx);
                        tma_load_3d(smem + s * STAGE_BYTES + A_BYTES, &maps.b[l], &full[s], kb * BK, tn * Ly.BN, 0);      // [plane][BN][64]
                    }
                    if (Ly.dep >= 0) {                   // wait until every column tile of the producing layer has published row tile tm
                        const int need = g.L[Ly.dep].tiles_n;
                        const int* c = g.counters + Ly.dep * g.tiles_m + tm;
                        while (ld_acquire(c) < need) __nanosleep(32);
                        asm volatile("fence.proxy.async;" ::: "memory");    // generic-proxy acquire -> async-proxy (TMA) reads
                    }
                    for (int kb = 0; kb < num_kb; ++kb, ++it) {
                        const int s = it % STAGES, k0 = kb * BK;
                        unsigned char* st = smem + s * STAGE_BYTES;
                        if (kb >= pre) {
                            mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                            mbar_expect_tx(&full[s], tx);
                            tma_load_3d(st + A_BYTES, &maps.b[l], &full[s], k0, tn * Ly.BN, 0);
                        }
                        tma_load_3d(st, &maps.a[l], &full[s], k0, tm * BM, 0);                   // [plane][128][64]
                    }
                }
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0, item = 0;
            for (int l = 0; l < g.n_layers; ++l) {
                const Chain16Layer& Ly = g.L[l];
                const int items = g.tiles_m * Ly.tiles_n, num_kb = (Ly.K + BK - 1) / BK;
                const uint32_t idesc = make_idesc(Ly.BN);
                const uint32_t b_plane = (uint32_t)Ly.BN * BK * 2u;
                for (int w = (blockIdx.x + Ly.rot) % G; w < items; w += G, ++item) {
                    const int acc_stage = item & 1;
                    mbar_wait(&tmem_empty[acc_stage], ((item >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc_stage * 128);
                    for (int kb = 0; kb < num_kb; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            const uint64_t a_hi = make_desc(a0 + kk * 32);
                            const uint64_t a_lo = make_desc(a0 + A_PLANE + kk * 32);
                            const uint64_t b_hi = make_desc(b0 + kk * 32);
                            const uint64_t b_lo = make_desc(b0 + b_plane + kk * 32);
                            umma_bf16(tmem_d, a_lo, b_hi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);   // kind::f16 (formats in idesc); small terms first
                            umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
                            umma_bf16(tmem_d, a_hi, b_hi, idesc, 1u);
                        }
                        umma_commit(&empty[s]);
                    }
                    umma_commit(&tmem_full[acc_stage]);
                }
            }
        }
    } else {
        // ===== epilogue: warp w <-> TMEM lanes 32 * (w % 4) .. +31, 32-column chunks c with c % 2 == w / 4 =====
        const int q = warp & 3, half = warp >> 2;
        unsigned char* stg = epi_stage + warp * EPI_STAGE_WARP;
        int item = 0;
        for (int l = 0; l < g.n_layers; ++l) {
            const Chain16Layer& Ly = g.L[l];
            const int items = g.tiles_m * Ly.tiles_n;
            for (int w = (blockIdx.x + Ly.rot) % G; w < items; w += G, ++item) {
                const int tm = w / Ly.tiles_n, tn = w - tm * Ly.tiles_n;
                const int acc_stage = item & 1;
                mbar_wait(&tmem_full[acc_stage], (item >> 1) & 1);
                tc_fence_after();
                const int row0 = tm * BM + q * 32, row = row0 + lane;
                const bool row_ok = row < g.M;
                for (int c0 = half * 32; c0 < Ly.BN; c0 += 64) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_stage * 128 + c0), v);
                    const int col0 = tn * Ly.BN + c0;
                    if (col0 >= Ly.N) continue;                                 // warp-uniform
                    const int nvalid = min(32, Ly.N - col0);
                    // bias: warp-uniform 128-bit loads (one L1 broadcast each), never shuffles
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(Ly.bias + col0) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(Ly.bias + col0 + j));
                            v[j] = fmaf(v[j], kWScaleInv, b4.x); v[j + 1] = fmaf(v[j + 1], kWScaleInv, b4.y);
                            v[j + 2] = fmaf(v[j + 2], kWScaleInv, b4.z); v[j + 3] = fmaf(v[j + 3], kWScaleInv, b4.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (j < nvalid) ? fmaf(v[j], kWScaleInv, __ldg(Ly.bias + col0 + j)) : 0.0f;
                    }
                    if (Ly.epi == CH_BIAS_ELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = elu_fp32(v[j]);
                        // the 32 x 32 block leaves as fp16 hi / lo planes through shared memory and ONE TMA store (rows >= M and
                        // columns >= N clipped by the map); lane = row, 16-byte chunk c of a row at chunk c ^ ((row >> 1) & 3)
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                        const int sw = (lane >> 1) & 3;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float x0 = v[8 * c + 2 * u], x1 = v[8 * c + 2 * u + 1];
                                ph[u] = pack_f16x2(x0, x1);
                                const float2 h = unpack_f16x2(ph[u]);
                                pl[u] = pack_f16x2(x0 - h.x, x1 - h.y);
                            }
                            unsigned char* d = stg + lane * 64 + ((c ^ sw) << 4);
                            *reinterpret_cast<uint4*>(d) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(d + 32 * 64) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&maps.c[l], stg, col0, row0, 0);
                            tma_store_commit();
                        }
                        continue;
                    }
                    if (!row_ok) continue;
                    float* dst = Ly.C + (int64_t)row * Ly.ldc + col0;
                    if (Ly.epi == CH_BIAS_SAMPLE) {
                        // ActorCritic.act + get_actions_log_prob (actor_critic.py:111-120) on the row this thread owns
                        const uint64_t stp = g.step_dev ? *g.step_dev : g.step;
                        float lp = 0.0f;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) {
                                const float mu = v[j];
                                const float sg = mu * 0.0f + __ldg(g.stdv + j);
                                float z;
                                if (g.eps) z = g.eps[(size_t)row * Ly.N + j];
                                else {
                                    HgPhilox r = hg_philox(g.seed, (uint32_t)row, (uint32_t)stp, HG_RNG_SAMPLE | ((uint32_t)(stp >> 32) << 8), j);
                                    z = hg_normal(r.c[0], r.c[1]);
                                }
                                const float a = mu + sg * z;
                                const float d = a - mu;
                                lp += -(d * d) / (2.0f * (sg * sg)) - logf(sg) - kLogSqrt2Pi;
                                dst[j] = mu;
                                g.actions[(size_t)row * Ly.N + j] = a;
                                g.sigma[(size_t)row * Ly.N + j] = sg;
                            }
                        g.logp[row] = lp;
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) dst[j] = v[j];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&tmem_empty[acc_stage]);                        // accumulator free for item + 2
                    // publish the tile: this warp's TMA stores have been written, and (generic stores of the output layers) every
                    // thread's stores are device-visible, before the counter moves
                    tma_store_wait_all();
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (threadIdx.x == 0) atomicAdd(g.counters + l * g.tiles_m + tm, 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
    // the last CTA to get here re-zeroes the counters (every wait of this launch has been satisfied by then)
    if (threadIdx.x == 0) {
        __threadfence();
        int* done = g.counters + MAX_CHAIN * g.tiles_m;
        if (atomicAdd(done, 1) == G - 1) {
            for (int i = 0; i < g.n_layers * g.tiles_m; ++i) g.counters[i] = 0;
            *done = 0;
            __threadfence();
        }
    }
}

// fp32 -> fp16 hi / lo planes of (x * scale); up to two tensors per launch (obs and critic obs).  One thread = 8 consecutive
// columns of a row: two 16-byte loads when the source row allows it, one 16-byte store per plane; pad columns get zeros.
struct SplitJob { const float* src; int64_t ld_src; uint16_t* dst; int64_t ld_dst, plane, rows, cols; };
__global__ void split_f16_kernel(SplitJob j0, SplitJob j1, float scale) {
    const SplitJob& j = blockIdx.y == 0 ? j0 : j1;
    const uint32_t groups = (uint32_t)(j.ld_dst >> 3);      // ld_dst % 8 == 0
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)j.rows * groups) return;
    const uint32_t r = (uint32_t)(i / groups), c = (uint32_t)(i - (uint64_t)r * groups) * 8u;
    const float* s = j.src + (int64_t)r * j.ld_src + c;
    float x[8];
    if (c + 8 <= (uint32_t)j.cols && ((reinterpret_cast<uintptr_t>(s) & 15u) == 0)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (c + k < (uint32_t)j.cols) ? __ldg(s + k) : 0.0f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x0 = x[2 * k] * scale, x1 = x[2 * k + 1] * scale;
        h[k] = pack_f16x2(x0, x1);
        const float2 hf = unpack_f16x2(h[k]);
        l[k] = pack_f16x2(x0 - hf.x, x1 - hf.y);
    }
    uint16_t* d = j.dst + (int64_t)r * j.ld_dst + c;
    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(d + j.plane) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- host ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int32_t load_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return hg_fail(HG_E_STATE, "cuTensorMapEncodeTiled unavailable");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}
// 3-D fp16 map over split planes {inner, rows of pitch ld, 2 planes}: box {64, box_rows, 2} SWIZZLE_128B (loads) or {32, 32, 2} SWIZZLE_64B (stores)
int32_t make_map(CUtensorMap* map, const uint16_t* base, uint64_t inner, uint64_t outer, uint64_t ld, uint64_t plane, uint32_t box_inner, uint32_t box_outer) {
    cuuint64_t dims[3] = {inner, outer, 2};
    cuuint64_t strides[2] = {ld * sizeof(uint16_t), plane * sizeof(uint16_t)};
    cuuint32_t box[3] = {box_inner, box_outer, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_hg_err, sizeof(g_hg_err), "hg_actor_critic_forward_f16: cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu plane=%llu",
                 (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, (unsigned long long)plane);
        return HG_E_ARG;
    }
    return 0;
}

struct Chain16Key {
    const void* p[9]; int64_t v[4];
    bool operator==(const Chain16Key& o) const { return memcmp(this, &o, sizeof(Chain16Key)) == 0; }
};
struct Chain16KeyHash {
    size_t operator()(const Chain16Key& k) const {
        uint64_t h = 1469598103934665603ull;
        const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
        for (size_t i = 0; i < sizeof(Chain16Key); ++i) { h ^= b[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
struct Chain16Plan { Chain16Maps maps; Chain16Args args; SplitJob in[2]; int n_in; int grid; };
std::unordered_map<Chain16Key, Chain16Plan, Chain16KeyHash> g_plans16;
std::mutex g_plans16_mu;

int64_t pad8(int64_t x) { return (x + 7) / 8 * 8; }
// uint16 elements of one net's planes: input + every hidden activation, each [2][M][pad8(width)]
int64_t net_scratch(const HgMlpDesc* net, int64_t M) {
    int64_t n = 0;
    for (int l = 0; l < net->n_layers; ++l) n += 2 * M * pad8(net->dims[l]);
    return n;
}

}  // namespace

extern "C" float hg_f16_weight_scale(void) { return kWScale; }

extern "C" int32_t hg_split_f16(const float* src, int64_t ld_src, const HgSplit* dst, int64_t rows, int64_t cols, float scale, void* stream) {
    HG_REQUIRE(src); HG_REQUIRE(dst); HG_REQUIRE(dst->p);
    if (rows <= 0 || cols <= 0 || ld_src < cols || dst->ld < cols || (dst->ld & 7) || (dst->plane & 7) || !hg_aligned16(dst->p))
        return hg_fail(HG_E_SIZE, "hg_split_f16: bad extents (ld >= cols, ld % 8 == 0, plane % 8 == 0, 16-byte aligned planes)");
    SplitJob j{src, ld_src, dst->p, dst->ld, dst->plane, rows, cols};
    const int64_t n = rows * (dst->ld >> 3);
    split_f16_kernel<<<dim3((unsigned)((n + 255) / 256), 1), 256, 0, (cudaStream_t)stream>>>(j, j, scale);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_split_f16");
}

extern "C" int64_t hg_actor_critic_f16_scratch_elems(const HgMlpDesc* actor, const HgMlpDesc* critic, int64_t M) {
    return (actor ? net_scratch(actor, M) : 0) + (critic ? net_scratch(critic, M) : 0);
}

extern "C" int32_t hg_actor_critic_forward_f16(const HgMlpDesc* actor, const HgMlpDesc* critic, const float* params, const uint16_t* w16,
                                               int64_t w16_plane, const float* obs, int64_t ld_obs, const float* cobs, int64_t ld_cobs,
                                               uint16_t* scratch16, float* mu, float* value, const HgMlpFwdOpts* sample, int32_t* counters,
                                               int64_t M, void* stream) {
    HG_REQUIRE(params); HG_REQUIRE(w16); HG_REQUIRE(counters); HG_REQUIRE(scratch16);
    if (!actor && !critic) return hg_fail(HG_E_NULL, "hg_actor_critic_forward_f16: actor and critic are both NULL");
    if (actor) { HG_REQUIRE(obs); HG_REQUIRE(mu); }
    if (critic) { HG_REQUIRE(cobs); HG_REQUIRE(value); }
    if (M <= 0 || M > (1 << 24)) return hg_fail(HG_E_SIZE, "hg_actor_critic_forward_f16: bad M");
    const int La = actor ? actor->n_layers : 0, Lc = critic ? critic->n_layers : 0;
    if ((actor && La < 1) || (critic && Lc < 1) || La + Lc > MAX_CHAIN) return hg_fail(HG_E_ALIGN, "hg_actor_critic_forward_f16: more than 8 layers in total");
    const bool want_sample = sample && sample->actions;
    if (want_sample && (!sample->std || !sample->log_prob || !sample->sigma)) return hg_fail(HG_E_NULL, "hg_actor_critic_forward_f16: sampling outputs are NULL");
    if (want_sample && (!actor || actor->dims[La] > 32)) return hg_fail(HG_E_ALIGN, "hg_actor_critic_forward_f16: sampling needs an actor with <= 32 actions");
    if (!hg_aligned16(w16) || (w16_plane & 7) || !hg_aligned16(scratch16)) return hg_fail(HG_E_ALIGN, "hg_actor_critic_forward_f16: w16 / scratch16 need 16-byte alignment, w16_plane % 8 == 0");
    if (int32_t rc = load_encode()) return rc;
    cudaStream_t st = (cudaStream_t)stream;

    Chain16Key key{};
    const void* ptrs[9] = {actor, critic, params, w16, obs, cobs, scratch16, mu, value};
    memcpy(key.p, ptrs, sizeof(ptrs));
    key.v[0] = ld_obs; key.v[1] = ld_cobs; key.v[2] = M; key.v[3] = (int64_t)(uintptr_t)counters ^ (w16_plane << 1);
    Chain16Plan plan;
    bool have = false;
    {
        std::lock_guard<std::mutex> lk(g_plans16_mu);
        auto it = g_plans16.find(key);
        if (it != g_plans16.end()) { plan = it->second; have = true; }
    }
    if (!have) {
        memset(&plan, 0, sizeof(plan));
        Chain16Args& g = plan.args;
        g.M = (int)M; g.tiles_m = (int)((M + BM - 1) / BM); g.counters = counters;
        const HgMlpDesc* nets[2] = {actor, critic};
        const float* X[2] = {obs, cobs};
        const int64_t ldx[2] = {ld_obs, ld_cobs};
        float* out[2] = {mu, value};
        uint16_t* base[2] = {scratch16, scratch16 + (actor ? net_scratch(actor, M) : 0)};
        const int Lmax = La > Lc ? La : Lc;
        int prev[2] = {-1, -1};
        uint16_t* cur[2] = {base[0], base[1]};          // planes of the current layer's INPUT
        int n = 0, rot = 0, max_items = 0;
        for (int which = 0; which < 2; ++which)
            if (nets[which]) {
                const int64_t K0 = nets[which]->dims[0], ld0 = pad8(K0);
                if (ldx[which] < K0) return hg_fail(HG_E_SIZE, "hg_actor_critic_forward_f16: input pitch < width");
                plan.in[plan.n_in++] = SplitJob{X[which], ldx[which], cur[which], ld0, M * ld0, M, K0};
            }
        for (int l = 0; l < Lmax; ++l)
            for (int which = 0; which < 2; ++which) {
                const HgMlpDesc* net = nets[which];
                if (!net || l >= net->n_layers) continue;
                const int K = net->dims[l], N = net->dims[l + 1];
                const bool last = (l + 1 == net->n_layers);
                const int64_t ld_in = pad8(K), ld_out = pad8(N);
                const uint16_t* in = cur[which];
                uint16_t* outp = cur[which] + 2 * M * ld_in;
                if ((net->ldw[l] & 7) || (net->w_off[l] & 7)) return hg_fail(HG_E_ALIGN, "hg_actor_critic_forward_f16: weight rows need pitch % 8 == 0");
                Chain16Layer& Ly = g.L[n];
                Ly.N = N; Ly.K = K;
                // tile width: enough column tiles that a layer's items cover the chip (same rule as hg_actor_critic_forward)
                int bn = ((N + 31) / 32) * 32;
                if (bn > BN_MAX) bn = BN_MAX;
                while (bn > 32 && g.tiles_m * ((N + bn - 1) / bn) < HG_NUM_SMS * 3 / 4 && (bn / 2) % 32 == 0) bn /= 2;
                Ly.BN = bn; Ly.tiles_n = (N + bn - 1) / bn;
                Ly.bias = params + net->b_off[l];
                Ly.C = last ? out[which] : nullptr;
                Ly.ldc = N;
                Ly.epi = last ? ((which == 0 && N <= 32 && Ly.tiles_n == 1) ? CH_BIAS_SAMPLE : CH_BIAS) : CH_BIAS_ELU;
                Ly.dep = prev[which];
                Ly.rot = rot;
                const int items = g.tiles_m * Ly.tiles_n;
                rot = (rot + items) % HG_NUM_SMS;
                if (items > max_items) max_items = items;
                if (int32_t rc = make_map(&plan.maps.a[n], in, K, M, ld_in, M * ld_in, 64, BM)) return rc;
                if (int32_t rc = make_map(&plan.maps.b[n], w16 + net->w_off[l], K, N, net->ldw[l], w16_plane, 64, bn)) return rc;
                if (!last) {
                    if (int32_t rc = make_map(&plan.maps.c[n], outp, N, M, ld_out, M * ld_out, 32, 32)) return rc;
                } else plan.maps.c[n] = plan.maps.a[n];
                prev[which] = n;
                cur[which] = outp;
                ++n;
            }
        g.n_layers = n;
        plan.grid = max_items < HG_NUM_SMS ? max_items : HG_NUM_SMS;
        for (int i = 0; i < n; ++i) g.L[i].rot %= plan.grid;
        std::lock_guard<std::mutex> lk(g_plans16_mu);
        if (g_plans16.size() > 256) g_plans16.clear();
        g_plans16.emplace(key, plan);
    }
    Chain16Args& g = plan.args;
    if (want_sample) {
        bool can = false;
        for (int i = 0; i < g.n_layers; ++i) can = can || g.L[i].epi == CH_BIAS_SAMPLE;
        if (!can) return hg_fail(HG_E_ALIGN, "hg_actor_critic_forward_f16: the actor's output layer cannot host the sampling epilogue");
        g.stdv = sample->std; g.eps = sample->eps; g.actions = sample->actions; g.logp = sample->log_prob; g.sigma = sample->sigma;
        g.seed = sample->seed; g.step = sample->step; g.step_dev = sample->step_dev;
    } else {
        for (int i = 0; i < g.n_layers; ++i)
            if (g.L[i].epi == CH_BIAS_SAMPLE) g.L[i].epi = CH_BIAS;
    }
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(mlp_chain16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return hg_fail((int32_t)e, cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    {   // network inputs -> fp16 planes (both nets in one launch)
        int64_t nmax = 0;
        for (int i = 0; i < plan.n_in; ++i) { const int64_t n = plan.in[i].rows * (plan.in[i].ld_dst >> 3); if (n > nmax) nmax = n; }
        split_f16_kernel<<<dim3((unsigned)((nmax + 255) / 256), plan.n_in), 256, 0, st>>>(plan.in[0], plan.in[plan.n_in - 1], 1.0f);
    }
    mlp_chain16_kernel<<<plan.grid, THREADS, SMEM_BYTES, st>>>(plan.maps, g);
    HG_LAUNCHED(2);
    return hg_cuda_status("hg_actor_critic_forward_f16");
}
