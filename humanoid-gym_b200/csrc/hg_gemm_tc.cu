// tcgen05 / TMEM / TMA GEMM for the ActorCritic MLP (sm_100a), fp32 in / fp32 out with 3xTF32
// split-compensation:   D = A_hi B_hi + A_lo B_hi + A_hi B_lo,   x_hi = tf32(x) (top 19 bits),  x_lo = x - x_hi
// which restores ~fp32 accuracy (gradient rel. error ~7e-7) while every multiply runs
// on the 5th-generation tensor cores.  `passes = 1` gives plain TF32.
//
// One kernel covers the three products of a Linear layer, without materialising any transpose:
//     forward  Y  = X  W^T   A = X  (K-major)    B = W  (K-major)
//     dgrad    dX = dZ W     A = dZ (K-major)    B = W  (MN-major: the GEMM's N index is W's contiguous dim)
//     wgrad    dW = dZ^T X   A = dZ (MN-major)   B = X  (MN-major), split-K over the batch, fp32 atomics
// (UMMA instruction-descriptor bits 15/16 select K- vs MN-major per operand.)
//
// CTA = 128 x BN output tile, 14 warps, warp-specialised:
//     warps 0-3  epilogue     tcgen05.ld 32 TMEM lanes each -> bias / ELU / ELU' / atomics -> global
//     warp  4    TMA producer cp.async.bulk.tensor (128B-swizzled 128x32 fp32 tiles, OOB zero fill)
//     warp  5    MMA issuer   one lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8), accumulator in TMEM
//     warps 6-13 splitter     raw fp32 tile -> lo tile (x - tf32(x)), elementwise so the swizzle is untouched
// 4-deep ring of raw {A, B} tile pairs (TMA look-ahead) + 2-deep ring of {A_lo, B_lo} pairs, mbarrier-linked; CTAs are
// persistent (one per SM) and the accumulator is double-buffered in TMEM so that epilogue and main loop overlap.
#include <cuda.h>

#include "hg_common.cuh"
#include "hg_tc_ptx.cuh"

namespace {

constexpr int BM = 128, BK = 32;
constexpr int RAW_STAGES = 4;                        // TMA look-ahead: raw {A, B} tiles
constexpr int LO_STAGES = 2;                         // splitter -> MMA: {A_lo, B_lo} tiles
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KB: 128 rows (or 4 MN-boxes) x 128 B
constexpr int PAIR_BYTES = 2 * TILE_BYTES;           // one {A, B} pair
constexpr int SPLIT_WARPS = 8;
constexpr int TC_THREADS = (6 + SPLIT_WARPS) * 32;   // 4 epilogue + TMA + MMA + splitters
constexpr int SMEM_BYTES = (RAW_STAGES + LO_STAGES) * PAIR_BYTES + 1024 /*align*/ + 256 /*barriers*/;

enum { EPI_STORE = 0, EPI_BIAS = 1, EPI_BIAS_ELU = 2, EPI_MUL_DELU = 3, EPI_ATOMIC = 4 };

struct TcArgs {
    float* C; const float* bias; const float* H;
    int M, N, K;                     // GEMM extents: C is M x N, reduction over K
    int64_t ldc, ldh;
    int BN;                          // 32 / 64 / 96 / 128
    int a_mn, b_mn;                  // operand majorness (0 = K-major, 1 = MN-major)
    int epi, passes;
    int kb_per_split;                // k-blocks (of 32) per split
    int splits;
    int hi_in_place;                 // 1: splitter rewrites the raw tile with its tf32 truncation
};

using namespace hgtc;

// UMMA shared-memory matrix descriptor, descriptor version 1 (sm_100).
//   K-major  tile [rows][32 fp32], SWIZZLE_128B (type 2)          : SBO = 1024 B (8 rows x 128 B), LBO unused (1)
//   MN-major tile 4 boxes of [32 k][32 fp32], 32-bit operands must use SWIZZLE_128B_BASE32B (type 1: 32-byte
//            chunks swizzled within 128 B, pattern period 4 rows)  : SBO = 512 B (4 k-rows), LBO = 4096 B (next 32 along MN)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(mn_major ? (4096 >> 4) : 1) << 16;
    d |= (uint64_t)((mn_major ? 512 : 1024) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(mn_major ? 1 : 2) << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format F32
    d |= 2u << 7;                       // a_format TF32
    d |= 2u << 10;                      // b_format TF32
    d |= (a_mn ? 1u : 0u) << 15;
    d |= (b_mn ? 1u : 0u) << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}

// Work item w = (m_tile, n_tile, split).  CTAs are persistent: CTA c processes items c, c + gridDim.x, ...;
// the {A,B,A_lo,B_lo} stage ring keeps rolling across items and the accumulator is double-buffered in TMEM
// (2 x 128 columns), so the epilogue of item i overlaps the main loop of item i + 1.
struct Work { int m0, n0, kb_begin, num_kb; };
__device__ __forceinline__ Work decode_work(const TcArgs& g, int w, int tiles_n, int tiles_mn, int num_kb_total) {
    Work r;
    const int split = w / tiles_mn, t = w - split * tiles_mn;
    const int tm = t / tiles_n, tn = t - tm * tiles_n;
    r.m0 = tm * BM;
    r.n0 = tn * g.BN;
    r.kb_begin = split * g.kb_per_split;
    r.num_kb = min(num_kb_total, r.kb_begin + g.kb_per_split) - r.kb_begin;
    return r;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_lo = smem + RAW_STAGES * PAIR_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_lo + LO_STAGES * PAIR_BYTES);
    uint64_t* full = bars;                          // [RAW] TMA -> splitter (and MMA)
    uint64_t* empty = full + RAW_STAGES;            // [RAW] MMA -> TMA
    uint64_t* ready = empty + RAW_STAGES;           // [LO]  splitter -> MMA
    uint64_t* lo_empty = ready + LO_STAGES;         // [LO]  MMA -> splitter
    uint64_t* tmem_full = lo_empty + LO_STAGES;     // [2]   MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;           // [2]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb_total = (g.K + BK - 1) / BK;
    const int tiles_n = (g.N + g.BN - 1) / g.BN, tiles_m = (g.M + BM - 1) / BM;
    const int tiles_mn = tiles_n * tiles_m;
    const int total_work = tiles_mn * g.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RAW_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < LO_STAGES; ++s) {
            mbar_init(&ready[s], SPLIT_WARPS);
            mbar_init(&lo_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            const uint32_t tx = TILE_BYTES + (uint32_t)g.BN * BK * 4;
            int it = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
                for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                    const int s = it % RAW_STAGES, k0 = (wk.kb_begin + kb) * BK;
                    mbar_wait(&empty[s], ((it / RAW_STAGES) & 1) ^ 1);
                    unsigned char* st = smem + s * PAIR_BYTES;
                    mbar_expect_tx(&full[s], tx);
                    if (!g.a_mn) tma_load_2d(st, &tmA, &full[s], k0, wk.m0);
                    else
                        for (int j = 0; j < BM / 32; ++j) tma_load_2d(st + j * 4096, &tmA, &full[s], wk.m0 + 32 * j, k0);
                    if (!g.b_mn) tma_load_2d(st + TILE_BYTES, &tmB, &full[s], k0, wk.n0);
                    else
                        for (int j = 0; j < g.BN / 32; ++j) tma_load_2d(st + TILE_BYTES + j * 4096, &tmB, &full[s], wk.n0 + 32 * j, k0);
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(g.BN, g.a_mn, g.b_mn);
            const uint32_t kstep_a = g.a_mn ? 1024 : 32, kstep_b = g.b_mn ? 1024 : 32;   // bytes per K = 8
            int it = 0, item = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++item) {
                const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
                const int acc_stage = item & 1;
                mbar_wait(&tmem_empty[acc_stage], ((item >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc_stage * 128);
                for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                    const int s = it % RAW_STAGES, l = it % LO_STAGES;
                    mbar_wait(&ready[l], (it / LO_STAGES) & 1);                   // lo tiles written (raw tiles landed before that)
                    mbar_wait(&full[s], (it / RAW_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + s * PAIR_BYTES), base_lo = smem_u32(smem_lo + l * PAIR_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        const uint64_t a_hi = make_desc(base + kk * kstep_a, g.a_mn);
                        const uint64_t b_hi = make_desc(base + TILE_BYTES + kk * kstep_b, g.b_mn);
                        const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
                        if (g.passes == 3) {
                            const uint64_t a_lo = make_desc(base_lo + kk * kstep_a, g.a_mn);
                            const uint64_t b_lo = make_desc(base_lo + TILE_BYTES + kk * kstep_b, g.b_mn);
                            umma_tf32(tmem_d, a_lo, b_hi, idesc, acc);          // small terms first
                            umma_tf32(tmem_d, a_hi, b_lo, idesc, 1u);
                            umma_tf32(tmem_d, a_hi, b_hi, idesc, 1u);
                        } else {
                            umma_tf32(tmem_d, a_hi, b_hi, idesc, acc);
                        }
                    }
                    umma_commit(&empty[s]);                                     // stages reusable once these MMAs retire
                    umma_commit(&lo_empty[l]);
                }
                umma_commit(&tmem_full[acc_stage]);
            }
        }
    } else if (warp >= 6) {
        // ===== splitter: x -> (tf32(x), x - tf32(x)), elementwise (swizzle-agnostic) =====
        const int t = threadIdx.x - 6 * 32;                                     // 0 .. 32*SPLIT_WARPS-1
        constexpr int NT = 32 * SPLIT_WARPS;
        const int nB4 = g.BN * BK / 4;                                          // float4 count of the B tile
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
            for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                const int s = it % RAW_STAGES, l = it % LO_STAGES;
                mbar_wait(&lo_empty[l], ((it / LO_STAGES) & 1) ^ 1);            // MMA is done with this lo pair
                mbar_wait(&full[s], (it / RAW_STAGES) & 1);
                if (g.passes == 3) {
                    float4* a = reinterpret_cast<float4*>(smem + s * PAIR_BYTES);
                    float4* b = a + TILE_BYTES / 16;
                    float4* alo = reinterpret_cast<float4*>(smem_lo + l * PAIR_BYTES);
                    float4* blo = alo + TILE_BYTES / 16;
                    auto split = [&](float4* raw, float4* lo, int i) {
                        // hi = truncation (what the tensor core does to a raw fp32 operand anyway);
                        // lo = residual, rounded to nearest tf32 so that its own truncation error vanishes
                        float4 x = raw[i], h, l;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = rna_tf32(x.x - h.x);
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = rna_tf32(x.y - h.y);
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = rna_tf32(x.z - h.z);
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = rna_tf32(x.w - h.w);
                        lo[i] = l;
                        if (g.hi_in_place) raw[i] = h;
                    };
#pragma unroll 4
                    for (int i = t; i < TILE_BYTES / 16; i += NT) split(a, alo, i);
#pragma unroll 4
                    for (int i = t; i < nB4; i += NT) split(b, blo, i);
                    fence_proxy_async();                                        // generic-proxy writes -> tensor-core reads
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[l]);
            }
        }
    } else {
        // ===== epilogue (warps 0-3 <-> TMEM lanes 32*warp .. +31) =====
        int item = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++item) {
            const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
            const int acc_stage = item & 1;
            mbar_wait(&tmem_full[acc_stage], (item >> 1) & 1);
            tc_fence_after();
            const int row = wk.m0 + warp * 32 + lane;
            const bool row_ok = row < g.M;
            for (int c0 = 0; c0 < g.BN; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc_stage * 128 + c0), v);
                const int col0 = wk.n0 + c0;
                if (col0 >= g.N) continue;                                  // warp-uniform
                float* dst = g.C + (int64_t)row * g.ldc + col0;
                const int nvalid = min(32, g.N - col0);
                const bool vec_ok = (nvalid == 32);
                if (g.epi == EPI_BIAS || g.epi == EPI_BIAS_ELU) {
                    // one coalesced load of the 32 bias values of this chunk, then warp shuffles
                    const float bl = (lane < nvalid) ? __ldg(g.bias + col0 + lane) : 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = v[j] + __shfl_sync(0xffffffffu, bl, j);
                        // nn.ELU(alpha=1): exp(x) - 1 for x <= 0 (absolute error ~1e-7, same form as torch's CUDA kernel)
                        v[j] = (g.epi == EPI_BIAS_ELU) ? (x > 0.0f ? x : __expf(x) - 1.0f) : x;
                    }
                } else if (g.epi == EPI_MUL_DELU) {
                    const float* h = g.H + (int64_t)row * g.ldh + col0;
                    if (row_ok && vec_ok && ((reinterpret_cast<uintptr_t>(h) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 hv = __ldg(reinterpret_cast<const float4*>(h + j));
                            v[j] *= (hv.x > 0.0f) ? 1.0f : (hv.x + 1.0f);
                            v[j + 1] *= (hv.y > 0.0f) ? 1.0f : (hv.y + 1.0f);
                            v[j + 2] *= (hv.z > 0.0f) ? 1.0f : (hv.z + 1.0f);
                            v[j + 3] *= (hv.w > 0.0f) ? 1.0f : (hv.w + 1.0f);
                        }
                    } else if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) {
                                float hv = __ldg(h + j);
                                v[j] *= (hv > 0.0f) ? 1.0f : (hv + 1.0f);
                            }
                    }
                }
                if (!row_ok) continue;
                if (g.epi == EPI_ATOMIC) {
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) atomicAdd(dst + j, v[j]);
                    }
                } else if (vec_ok && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < nvalid) dst[j] = v[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc_stage]);                 // accumulator free for item + 2
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
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

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows of pitch `ld` elements; 128B-swizzled boxes
int32_t make_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                 bool mn_major) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_hg_err, sizeof(g_hg_err), "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
        return HG_E_ARG;
    }
    return 0;
}

}  // namespace

// C (M x N) = op(A) op(B) over K, see HgGemm in hg_b200.h
extern "C" int32_t hg_gemm_tf32(const HgGemm* d, void* stream) {
    HG_REQUIRE(d); HG_REQUIRE(d->A); HG_REQUIRE(d->B); HG_REQUIRE(d->C);
    if (d->M <= 0 || d->N <= 0 || d->K <= 0) return hg_fail(HG_E_SIZE, "hg_gemm_tf32: bad extents");
    if (d->passes != 1 && d->passes != 3) return hg_fail(HG_E_ARG, "hg_gemm_tf32: passes must be 1 (plain TF32) or 3 (3xTF32)");
    if ((d->lda & 3) || (d->ldb & 3) || !hg_aligned16(d->A) || !hg_aligned16(d->B))
        return hg_fail(HG_E_ALIGN, "hg_gemm_tf32: operands need 16-byte aligned base and row pitch (TMA)");
    if ((d->epilogue == EPI_BIAS || d->epilogue == EPI_BIAS_ELU) && !d->bias) return hg_fail(HG_E_NULL, "hg_gemm_tf32: bias is NULL");
    if (d->epilogue == EPI_MUL_DELU && !d->H) return hg_fail(HG_E_NULL, "hg_gemm_tf32: H is NULL");
    if (d->epilogue < 0 || d->epilogue > EPI_ATOMIC) return hg_fail(HG_E_ARG, "hg_gemm_tf32: bad epilogue");
    if (int32_t rc = load_encode()) return rc;
    cudaStream_t st = (cudaStream_t)stream;

    TcArgs g{};
    g.C = d->C; g.bias = d->bias; g.H = d->H;
    g.M = d->M; g.N = d->N; g.K = d->K;
    g.ldc = d->ldc; g.ldh = d->ldh;
    g.a_mn = d->a_mn_major ? 1 : 0; g.b_mn = d->b_mn_major ? 1 : 0;
    g.epi = d->epilogue; g.passes = d->passes;
    g.hi_in_place = d->trust_hw_truncation ? 0 : 1;
    int bn = ((d->N + 31) / 32) * 32;
    g.BN = bn > 128 ? 128 : bn;
    const int num_kb = (d->K + BK - 1) / BK;
    int splits = d->split_k > 0 ? d->split_k : 1;
    if (splits > num_kb) splits = num_kb;
    if (splits > 1 && d->epilogue != EPI_ATOMIC) return hg_fail(HG_E_ARG, "hg_gemm_tf32: split_k needs the atomic epilogue");
    g.kb_per_split = (num_kb + splits - 1) / splits;
    splits = (num_kb + g.kb_per_split - 1) / g.kb_per_split;

    // K-major operand: tensor (inner = K, outer = rows), box {32, rows}.  MN-major: (inner = rows, outer = K), box {32, 32}.
    CUtensorMap tmA, tmB;
    int32_t rc;
    if (!g.a_mn) rc = make_map(&tmA, d->A, d->K, d->M, d->lda, BK, BM, false);
    else rc = make_map(&tmA, d->A, d->M, d->K, d->lda, 32, BK, true);
    if (rc) return rc;
    if (!g.b_mn) rc = make_map(&tmB, d->B, d->K, d->N, d->ldb, BK, g.BN, false);
    else rc = make_map(&tmB, d->B, d->N, d->K, d->ldb, 32, BK, true);
    if (rc) return rc;

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return hg_fail((int32_t)e, cudaGetErrorString(e));
        attr_set = true;
    }
    g.splits = splits;
    const int total_work = ((d->N + g.BN - 1) / g.BN) * ((d->M + BM - 1) / BM) * splits;
    const int grid = total_work < HG_NUM_SMS ? total_work : HG_NUM_SMS;       // persistent: one CTA per SM
    gemm_tc_kernel<<<grid, TC_THREADS, SMEM_BYTES, st>>>(tmA, tmB, g);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_gemm_tf32");
}
