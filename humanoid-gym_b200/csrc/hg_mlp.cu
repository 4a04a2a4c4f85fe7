// ActorCritic MLP forward / backward (algo/ppo/actor_critic.py:54-77, autograd in ppo.py:170-173).
//
// This file holds the exact-fp32 CUDA-core GEMM path: C(i,j) = sum_p A(i,p) * B(p,j) with arbitrary
// element strides on both operands so that the three products of a Linear layer
//     forward  Y  = X  W^T        (A = X  row-major,   B(p,j) = W[j][p])
//     dgrad    dX = dZ W          (A = dZ row-major,   B(p,j) = W[p][j])
//     wgrad    dW = dZ^T X        (A(i,p) = dZ[p][i],  B(p,j) = X[p][j], split-K over the batch)
// share one kernel, with the bias+ELU and ELU' epilogues fused.  It is the bit-for-bit-fp32 reference
// for the tcgen05 path (hg_gemm_tc.cu) and the fallback for shapes the tensor-core tiles do not cover
// (K = 219 / 705 row pitches that TMA cannot address, N = 12 / 1 output layers).
#include <stdlib.h>

#include "hg_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;
constexpr int GEMM_THREADS = (BM / TM) * (BN / TN);   // 256

enum Epilogue : int { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_ELU = 2, EPI_MUL_DELU = 3, EPI_ATOMIC = 4 };

struct GemmArgs {
    const float* A; const float* B; float* C;
    const float* bias;      // (N) for EPI_BIAS*
    const float* h;         // (M, ldh) post-ELU activations for EPI_MUL_DELU
    int M, N, K;
    int64_t sai, sap, sbp, sbj, ldc, ldh;
    int k_chunk;            // split-K: p range per blockIdx.z
};

// A_IMAJ: A is contiguous along i (sai == 1); otherwise contiguous along p.
// B_JMAJ: B is contiguous along j (sbj == 1); otherwise contiguous along p.
template <int EPI, bool A_IMAJ, bool B_JMAJ>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + 4];   // +4: conflict-free transposed stores
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int p_begin = blockIdx.z * g.k_chunk;
    const int p_end = min(g.K, p_begin + g.k_chunk);

    // global->smem mapping: 4 elements per thread per operand
    int a_i[4], a_p[4], b_j[4], b_p[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int idx = tid + r * GEMM_THREADS;               // 0..1023
        if (A_IMAJ) { a_p[r] = idx / BM; a_i[r] = idx % BM; } else { a_i[r] = idx / BK; a_p[r] = idx % BK; }
        if (B_JMAJ) { b_p[r] = idx / BN; b_j[r] = idx % BN; } else { b_j[r] = idx / BK; b_p[r] = idx % BK; }
    }
    float ra[4], rb[4];
    auto fetch = [&](int p0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = i0 + a_i[r], p = p0 + a_p[r];
            ra[r] = (i < g.M && p < p_end) ? __ldg(g.A + (int64_t)i * g.sai + (int64_t)p * g.sap) : 0.0f;
            int j = j0 + b_j[r];
            p = p0 + b_p[r];
            rb[r] = (j < g.N && p < p_end) ? __ldg(g.B + (int64_t)p * g.sbp + (int64_t)j * g.sbj) : 0.0f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 4; ++r) { As[buf][a_p[r]][a_i[r]] = ra[r]; Bs[buf][b_p[r]][b_j[r]] = rb[r]; }
    };

    const int ty = tid / (BN / TN), tx = tid % (BN / TN);   // 16 x 16
    float acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = 0.0f;

    int buf = 0;
    fetch(p_begin);
    stash(0);
    __syncthreads();
    for (int p0 = p_begin; p0 < p_end; p0 += BK) {
        const bool more = p0 + BK < p_end;
        if (more) fetch(p0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < TM; ++a)
#pragma unroll
                for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

#pragma unroll
    for (int a = 0; a < TM; ++a) {
        int i = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + (a - 4));
        if (i >= g.M) continue;
#pragma unroll
        for (int b = 0; b < TN; ++b) {
            int j = j0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
            if (j >= g.N) continue;
            float v = acc[a][b];
            float* dst = g.C + (int64_t)i * g.ldc + j;
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_ELU) v += g.bias[j];
            if (EPI == EPI_BIAS_ELU) v = v > 0.0f ? v : expm1f(v);          // nn.ELU(alpha=1)
            if (EPI == EPI_MUL_DELU) {                                       // ELU'(z) from h = ELU(z)
                float h = g.h[(int64_t)i * g.ldh + j];
                v *= (h > 0.0f) ? 1.0f : (h + 1.0f);
            }
            if (EPI == EPI_ATOMIC) atomicAdd(dst, v); else *dst = v;
        }
    }
}

template <int EPI>
int32_t launch_gemm(const GemmArgs& g, int splits, cudaStream_t st) {
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splits);
    const bool a_imaj = (g.sai == 1 && g.sap != 1), b_jmaj = (g.sbj == 1);
    if (a_imaj && b_jmaj) gemm_kernel<EPI, true, true><<<grid, GEMM_THREADS, 0, st>>>(g);
    else if (a_imaj) gemm_kernel<EPI, true, false><<<grid, GEMM_THREADS, 0, st>>>(g);
    else if (b_jmaj) gemm_kernel<EPI, false, true><<<grid, GEMM_THREADS, 0, st>>>(g);
    else gemm_kernel<EPI, false, false><<<grid, GEMM_THREADS, 0, st>>>(g);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg gemm");
}

// db[n] = sum_m dZ[m][n].  Block = 32 columns x 8 row-lanes; each warp reads 128-byte row segments (coalesced),
// 4 independent loads in flight per thread; partial sums meet in shared memory, one atomicAdd per column per block.
constexpr int CS_ROWS = 2048;
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dZ, float* __restrict__ db, int M, int N) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const int m0 = blockIdx.y * CS_ROWS, m1 = min(M, m0 + CS_ROWS);
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (n < N) {
        int m = m0 + ty;
        for (; m + 24 < m1; m += 32) {
            s0 += __ldg(dZ + (int64_t)m * N + n);
            s1 += __ldg(dZ + (int64_t)(m + 8) * N + n);
            s2 += __ldg(dZ + (int64_t)(m + 16) * N + n);
            s3 += __ldg(dZ + (int64_t)(m + 24) * N + n);
        }
        for (; m < m1; m += 8) s0 += __ldg(dZ + (int64_t)m * N + n);
    }
    __shared__ float sh[8][33];
    sh[ty][tx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][tx];
        atomicAdd(db + n, t);
    }
}

// 128-bit variant (N % 4 == 0, 16-byte aligned dZ): a warp reads 512 contiguous bytes of a row, every thread keeps
// 8 independent 16-byte loads in flight (32 KB per block), so the kernel runs at HBM / L2 rate instead of at
// load latency.  Block = 128 columns x 512 rows; partial sums meet in shared memory, 128 atomicAdds per block.
constexpr int CS4_ROWS = 512;
__global__ void __launch_bounds__(256) colsum4_kernel(const float* __restrict__ dZ, float* __restrict__ db, int M, int N) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nc4 = N >> 2;
    const int c4 = blockIdx.x * 32 + lane;
    const int m0 = blockIdx.y * CS4_ROWS, m1 = min(M, m0 + CS4_ROWS);
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (c4 < nc4) {
        const float4* p = reinterpret_cast<const float4*>(dZ) + c4;
        int m = m0 + w;
        for (; m + 56 < m1; m += 64) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldcs(p + (int64_t)(m + 8 * i) * nc4);
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
        }
        for (; m < m1; m += 8) {
            float4 v = __ldcs(p + (int64_t)m * nc4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    __shared__ float4 sh[8][33];
    sh[w][lane] = acc;
    __syncthreads();
    if (w == 0 && c4 < nc4) {
        float4 t = sh[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) { t.x += sh[k][lane].x; t.y += sh[k][lane].y; t.z += sh[k][lane].z; t.w += sh[k][lane].w; }
        float* d = db + 4 * c4;
        atomicAdd(d, t.x); atomicAdd(d + 1, t.y); atomicAdd(d + 2, t.z); atomicAdd(d + 3, t.w);
    }
}

// ---- skinny layers (N_out <= 16: the 128 -> 12 / 128 -> 1 heads) -----------------------------------------------
// A 128-row MMA tile would be > 87 % padding here and the generic CUDA-core GEMM splits K into hundreds of
// atomically-merged slices; both backward products of such a layer are really one streaming pass over the
// (M, K) activation matrix, so they get their own HBM-rate kernels.
constexpr int SK_ROWS = 128;
// dW[n][k] += sum_{m in chunk} dZ[m][n] X[m][k];  grid (ceil(K/128), ceil(M/SK_ROWS)), block 128 (one k per thread)
template <int NO>
__global__ void __launch_bounds__(128) wgrad_skinny_kernel(const float* __restrict__ dZ, const float* __restrict__ X, int64_t ldx,
                                                           float* __restrict__ dW, int64_t ldw, int M, int N, int K) {
    __shared__ float dz[SK_ROWS][NO];
    const int m0 = blockIdx.y * SK_ROWS, rows = min(SK_ROWS, M - m0);
    for (int i = threadIdx.x; i < SK_ROWS * NO; i += 128) {
        int r = i / NO, n = i - r * NO;
        dz[r][n] = (r < rows && n < N) ? dZ[(int64_t)(m0 + r) * N + n] : 0.0f;
    }
    __syncthreads();
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= K) return;
    float acc[NO];
#pragma unroll
    for (int n = 0; n < NO; ++n) acc[n] = 0.0f;
    const float* xp = X + (int64_t)m0 * ldx + k;
    int r = 0;
    for (; r + 8 <= rows; r += 8) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __ldcs(xp + (int64_t)(r + i) * ldx);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int n = 0; n < NO; ++n) acc[n] += dz[r + i][n] * x[i];
    }
    for (; r < rows; ++r) {
        float x = __ldcs(xp + (int64_t)r * ldx);
#pragma unroll
        for (int n = 0; n < NO; ++n) acc[n] += dz[r][n] * x;
    }
#pragma unroll
    for (int n = 0; n < NO; ++n)
        if (n < N) atomicAdd(dW + (int64_t)n * ldw + k, acc[n]);
}

// dH[m][k] = (sum_n dZ[m][n] W[n][k]) * ELU'(h[m][k]);  one float4 of k per thread, K % 4 == 0, 16-byte aligned rows
template <int NO>
__global__ void __launch_bounds__(256) dgrad_skinny_kernel(const float* __restrict__ dZ, const float* __restrict__ W, int64_t ldw,
                                                           const float* __restrict__ H, float* __restrict__ dH, int M, int N, int K) {
    const int kq = K >> 2;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)M * kq) return;
    const int m = (int)(idx / kq), k4 = (int)(idx - (int64_t)m * kq);
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int n = 0; n < NO; ++n) {
        if (n < N) {
            const float d = __ldg(dZ + (int64_t)m * N + n);
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw) + k4);
            v.x += d * w.x; v.y += d * w.y; v.z += d * w.z; v.w += d * w.w;
        }
    }
    const float4 h = __ldcs(reinterpret_cast<const float4*>(H + (int64_t)m * K) + k4);
    v.x *= (h.x > 0.0f) ? 1.0f : (h.x + 1.0f);
    v.y *= (h.y > 0.0f) ? 1.0f : (h.y + 1.0f);
    v.z *= (h.z > 0.0f) ? 1.0f : (h.z + 1.0f);
    v.w *= (h.w > 0.0f) ? 1.0f : (h.w + 1.0f);
    reinterpret_cast<float4*>(dH + (int64_t)m * K)[k4] = v;
}

template <int NO>
void launch_wgrad_skinny(const float* dZ, const float* X, int64_t ldx, float* dW, int64_t ldw, int M, int N, int K, cudaStream_t st) {
    dim3 grid((K + 127) / 128, (M + SK_ROWS - 1) / SK_ROWS);
    wgrad_skinny_kernel<NO><<<grid, 128, 0, st>>>(dZ, X, ldx, dW, ldw, M, N, K);
}
template <int NO>
void launch_dgrad_skinny(const float* dZ, const float* W, int64_t ldw, const float* H, float* dH, int M, int N, int K, cudaStream_t st) {
    int64_t total = (int64_t)M * (K >> 2);
    dgrad_skinny_kernel<NO><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dZ, W, ldw, H, dH, M, N, K);
}

int g_gemm_mode = -1;   // -1: read HG_GEMM on first use
int gemm_mode() {
    if (g_gemm_mode < 0) {
        const char* e = getenv("HG_GEMM");
        g_gemm_mode = 4;                                                         // bf16x3 update + 3xTF32 rollout (hg_b200.h)
        if (e && (!strcmp(e, "simt") || !strcmp(e, "0"))) g_gemm_mode = 0;
        if (e && (!strcmp(e, "3xtf32") || !strcmp(e, "1"))) g_gemm_mode = 1;
        if (e && (!strcmp(e, "tf32") || !strcmp(e, "2"))) g_gemm_mode = 2;
    }
    return g_gemm_mode;
}

// Try the tensor-core kernel for C = A x B; returns HG_E_ALIGN-style "not eligible" as 1 so the caller falls back.
// a_mn / b_mn: operand is stored with its M / N index contiguous (see HgGemm).
int32_t try_tc(const float* A, int64_t lda, bool a_mn, const float* B, int64_t ldb, bool b_mn, float* C, int64_t ldc, int M, int N,
               int K, int epi, const float* bias, const float* H, int64_t ldh, int split_k, cudaStream_t st,
               const float* B_lo = nullptr, const HgMlpFwdOpts* sample = nullptr) {
    const int mode = gemm_mode();
    if (mode == 0) return 1;
    if ((lda & 3) || (ldb & 3) || !hg_aligned16(A) || !hg_aligned16(B)) return 1;
    HgGemm d{};
    d.A = A; d.B = B; d.C = C; d.bias = bias; d.H = H;
    d.M = M; d.N = N; d.K = K;
    d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.ldh = ldh;
    d.a_mn_major = a_mn; d.b_mn_major = b_mn;
    d.epilogue = epi; d.passes = (mode == 2) ? 1 : 3; d.split_k = split_k;
    d.trust_hw_truncation = 1;      // verified on B200: kind::tf32 ignores the low 13 mantissa bits (tests/test_gemm_tc_gpu.py)
    d.B_lo = (d.passes == 3) ? B_lo : nullptr;
    if (sample) {                   // output layer of the actor during the rollout: PPO.act fused into the epilogue
        d.epilogue = 5;
        d.sample_std = sample->std; d.sample_eps = sample->eps; d.sample_actions = sample->actions;
        d.sample_log_prob = sample->log_prob; d.sample_sigma = sample->sigma;
        d.sample_seed = sample->seed; d.sample_step = sample->step; d.sample_step_dev = sample->step_dev;
    }
    return hg_gemm_tf32(&d, (void*)st);
}

int32_t check_net(const HgMlpDesc* net) {
    HG_REQUIRE(net);
    if (net->n_layers < 1 || net->n_layers > HG_MAX_LAYERS) return hg_fail(HG_E_ARG, "HgMlpDesc: bad n_layers");
    for (int l = 0; l <= net->n_layers; ++l)
        if (net->dims[l] < 1) return hg_fail(HG_E_SIZE, "HgMlpDesc: bad layer width");
    for (int l = 0; l < net->n_layers; ++l)
        if (net->ldw[l] < net->dims[l]) return hg_fail(HG_E_SIZE, "HgMlpDesc: ldw < layer input width");
    return 0;
}

}  // namespace

extern "C" int32_t hg_set_gemm_mode(int32_t mode) {
    int prev = gemm_mode();
    if (mode == 0 || mode == 1 || mode == 2 || mode == 4) g_gemm_mode = mode;      // any other value: query only
    return prev;
}

__global__ void tf32_residual_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = src[i];
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - h));
    dst[i] = __uint_as_float(r);
}
extern "C" int32_t hg_tf32_residual(const float* src, float* dst, int64_t n, void* stream) {
    HG_REQUIRE(src); HG_REQUIRE(dst);
    if (n <= 0) return hg_fail(HG_E_SIZE, "hg_tf32_residual: bad n");
    tf32_residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_tf32_residual");
}

extern "C" int32_t hg_mlp_forward(const HgMlpDesc* net, const float* params, const float* X, int64_t ldx,
                                  float* hidden, float* out, int64_t M, void* stream) {
    return hg_mlp_forward_ex(net, params, X, ldx, hidden, out, M, nullptr, stream);
}

extern "C" int32_t hg_mlp_forward_ex(const HgMlpDesc* net, const float* params, const float* X, int64_t ldx,
                                     float* hidden, float* out, int64_t M, const HgMlpFwdOpts* opts, void* stream) {
    if (int32_t rc = check_net(net)) return rc;
    const float* params_lo = opts ? opts->params_lo : nullptr;
    const bool want_sample = opts && opts->actions;
    if (want_sample && (!opts->std || !opts->log_prob || !opts->sigma)) return hg_fail(HG_E_NULL, "hg_mlp_forward_ex: sampling outputs are NULL");
    HG_REQUIRE(params); HG_REQUIRE(X); HG_REQUIRE(out);
    if (net->n_layers > 1) HG_REQUIRE(hidden);
    if (M <= 0 || M > (1 << 28) || ldx < net->dims[0]) return hg_fail(HG_E_SIZE, "hg_mlp_forward: bad M/ldx");
    cudaStream_t st = (cudaStream_t)stream;
    const float* in = X;
    int64_t ld_in = ldx;
    float* h = hidden;
    bool sampled = false;
    for (int l = 0; l < net->n_layers; ++l) {
        const int K = net->dims[l], N = net->dims[l + 1];
        const bool last = (l + 1 == net->n_layers);
        const float* W = params + net->w_off[l];
        const int64_t ldw = net->ldw[l];
        float* dst = last ? out : h;
        const bool fuse_sample = last && want_sample && N <= 32 && gemm_mode() != 0;
        int32_t rc = try_tc(in, ld_in, false, W, ldw, false, dst, N, (int)M, N, K, last ? 1 : 2, params + net->b_off[l], nullptr, 0, 1, st,
                            params_lo ? params_lo + net->w_off[l] : nullptr, fuse_sample ? opts : nullptr);
        if (rc == 0 && fuse_sample) sampled = true;
        if (rc == 1) {
            GemmArgs g{};
            g.A = in; g.sai = ld_in; g.sap = 1;
            g.B = W; g.sbp = 1; g.sbj = ldw;
            g.C = dst; g.ldc = N;
            g.bias = params + net->b_off[l];
            g.M = (int)M; g.N = N; g.K = K; g.k_chunk = K;
            rc = last ? launch_gemm<EPI_BIAS>(g, 1, st) : launch_gemm<EPI_BIAS_ELU>(g, 1, st);
        }
        if (rc) return rc;
        in = h; ld_in = N;
        h += M * N;
    }
    if (want_sample && !sampled) {       // engines without the fused epilogue: the stand-alone kernel on the mean just written
        const int A = net->dims[net->n_layers];
        return hg_policy_sample(out, opts->std, opts->eps, opts->seed, opts->step, opts->step_dev, opts->actions, opts->log_prob,
                                opts->sigma, M, A, stream);
    }
    return 0;
}

extern "C" int32_t hg_mlp_backward(const HgMlpDesc* net, const float* params, const float* X, int64_t ldx,
                                   const float* hidden, const float* dY, float* dhidden, float* grads,
                                   int64_t M, void* stream) {
    if (int32_t rc = check_net(net)) return rc;
    HG_REQUIRE(params); HG_REQUIRE(X); HG_REQUIRE(dY); HG_REQUIRE(grads);
    if (net->n_layers > 1) { HG_REQUIRE(hidden); HG_REQUIRE(dhidden); }
    if (M <= 0 || M > (1 << 28) || ldx < net->dims[0]) return hg_fail(HG_E_SIZE, "hg_mlp_backward: bad M/ldx");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = net->n_layers;
    // element offset of hidden layer l's output (l = 1..L-1) inside hidden / dhidden
    int64_t off[HG_MAX_LAYERS + 1];
    off[0] = 0; off[1] = 0;
    for (int l = 1; l < L; ++l) off[l + 1] = off[l] + M * net->dims[l];
    // zero this net's gradients: one memset when its blocks are laid out back to back (w0 b0 w1 b1 ...), else per block
    bool contiguous = true;
    int64_t end = net->w_off[0];
    for (int l = 0; l < L && contiguous; ++l) {
        const int64_t n_out = net->dims[l + 1];
        contiguous = net->w_off[l] == end && net->b_off[l] == net->w_off[l] + n_out * net->ldw[l];
        end = net->b_off[l] + (n_out + 3) / 4 * 4;
    }
    if (contiguous) {
        // the last bias block is zeroed to its exact length (what follows it may belong to somebody else)
        const int64_t last = net->b_off[L - 1] + net->dims[L];
        cudaMemsetAsync(grads + net->w_off[0], 0, sizeof(float) * (size_t)(last - net->w_off[0]), st);
    }
    const float* dZ = dY;                                   // gradient w.r.t. layer l's pre-activation
    for (int l = L - 1; l >= 0; --l) {
        const int K = net->dims[l], N = net->dims[l + 1];
        const float* in = (l == 0) ? X : hidden + off[l];
        const int64_t ld_in = (l == 0) ? ldx : K;
        float* dW = grads + net->w_off[l];
        float* db = grads + net->b_off[l];
        const int64_t ldw = net->ldw[l];
        if (!contiguous) {
            cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * ldw, st);
            cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, st);
        }
        {   // dW[n][k] = sum_m dZ[m][n] X[m][k]
            int tiles = ((N + BM - 1) / BM) * ((K + BN - 1) / BN);
            int64_t want_tc = (2 * HG_NUM_SMS + tiles - 1) / tiles, cap_tc = (M + 511) / 512;
            int split_tc = (int)(want_tc < cap_tc ? want_tc : cap_tc);
            if (split_tc < 1) split_tc = 1;
            int32_t rc = try_tc(dZ, N, true, in, ld_in, true, dW, ldw, N, K, (int)M, 4, nullptr, nullptr, 0, split_tc, st);
            if (rc == 1 && N <= 16) {
                if (N == 1) launch_wgrad_skinny<1>(dZ, in, ld_in, dW, ldw, (int)M, N, K, st);
                else if (N <= 4) launch_wgrad_skinny<4>(dZ, in, ld_in, dW, ldw, (int)M, N, K, st);
                else if (N <= 8) launch_wgrad_skinny<8>(dZ, in, ld_in, dW, ldw, (int)M, N, K, st);
                else launch_wgrad_skinny<16>(dZ, in, ld_in, dW, ldw, (int)M, N, K, st);
                HG_LAUNCHED(1);
                rc = hg_cuda_status("hg wgrad (skinny)");
            } else if (rc == 1) {
                GemmArgs g{};
                g.A = dZ; g.sai = 1; g.sap = N;
                g.B = in; g.sbp = ld_in; g.sbj = 1;
                g.C = dW; g.ldc = ldw;
                g.M = N; g.N = K; g.K = (int)M;
                int64_t want = (4 * HG_NUM_SMS + tiles - 1) / tiles, cap = (M + 255) / 256;
                int splits = (int)(want < cap ? want : cap);
                if (splits < 1) splits = 1;
                g.k_chunk = (int)(((M + splits - 1) / splits + BK - 1) / BK * BK);
                splits = (int)((M + g.k_chunk - 1) / g.k_chunk);
                rc = launch_gemm<EPI_ATOMIC>(g, splits, st);
            }
            if (rc) return rc;
        }
        if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(dZ) & 15u) == 0) {
            dim3 grid((N / 4 + 31) / 32, (unsigned)((M + CS4_ROWS - 1) / CS4_ROWS));
            colsum4_kernel<<<grid, 256, 0, st>>>(dZ, db, (int)M, N);
            HG_LAUNCHED(1);
        } else {
            dim3 grid((N + 31) / 32, (unsigned)((M + CS_ROWS - 1) / CS_ROWS));
            colsum_kernel<<<grid, 256, 0, st>>>(dZ, db, (int)M, N);
            HG_LAUNCHED(1);
        }
        if (l > 0) {   // dZ_{l-1} = (dZ_l W_l) * ELU'(h_{l-1})
            const float* W = params + net->w_off[l];
            int32_t rc = try_tc(dZ, N, false, W, ldw, true, dhidden + off[l], K, (int)M, K, N, 3, nullptr, hidden + off[l], K, 1, st);
            if (rc == 1 && N <= 16 && (K & 3) == 0 && (ldw & 3) == 0 && hg_aligned16(W) && hg_aligned16(hidden + off[l]) &&
                hg_aligned16(dhidden + off[l])) {
                float* dH = dhidden + off[l];
                const float* H = hidden + off[l];
                if (N == 1) launch_dgrad_skinny<1>(dZ, W, ldw, H, dH, (int)M, N, K, st);
                else if (N <= 4) launch_dgrad_skinny<4>(dZ, W, ldw, H, dH, (int)M, N, K, st);
                else if (N <= 8) launch_dgrad_skinny<8>(dZ, W, ldw, H, dH, (int)M, N, K, st);
                else launch_dgrad_skinny<16>(dZ, W, ldw, H, dH, (int)M, N, K, st);
                HG_LAUNCHED(1);
                rc = hg_cuda_status("hg dgrad (skinny)");
            } else if (rc == 1) {
                GemmArgs g{};
                g.A = dZ; g.sai = N; g.sap = 1;
                g.B = W; g.sbp = ldw; g.sbj = 1;
                g.C = dhidden + off[l]; g.ldc = K;
                g.h = hidden + off[l]; g.ldh = K;
                g.M = (int)M; g.N = K; g.K = N; g.k_chunk = N;
                rc = launch_gemm<EPI_MUL_DELU>(g, 1, st);
            }
            if (rc) return rc;
            dZ = dhidden + off[l];
        }
    }
    return hg_cuda_status("hg_mlp_backward");
}
