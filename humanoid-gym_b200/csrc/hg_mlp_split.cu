// ActorCritic MLP forward / backward on split-precision (bf16 hi/lo) tensors: the PPO.update path
// (algo/ppo/ppo.py:150-173 -> actor_critic.py:54-77 and its autograd).
//
// Hidden layers run on hg_gemm_bf16x3 (tcgen05, hg_gemm_bf3.cu); activations, their gradients and the weights exist
// only as split planes, written by the producing epilogue.  The <= 16-wide output layer (128 -> 12 / 128 -> 1) is
// 0.3 % of the FLOPs and > 87 % padding on a 128-row MMA tile, so it runs on CUDA cores as streaming passes over the
// last hidden activation:
//     head_forward_kernel   out = h W^T + b
//     head_backward_kernel  dW, db of the head, dZ of the last hidden layer (x ELU', split store) and that layer's
//                           bias gradient -- ONE pass over h instead of four.
#include "hg_common.cuh"


namespace {

constexpr int HEAD_MAX_N = 16;

__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float x0, float x1) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}

// out[m][n] = b[n] + sum_k h[m][k] W[n][k]; one thread per row, W^T staged in shared memory (broadcast reads).
// K % 8 == 0, h planes 16-byte aligned rows.
template <int NO>
__global__ void __launch_bounds__(128) head_forward_kernel(const uint16_t* __restrict__ hs, int64_t ldh, int64_t hplane,
                                                           const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias,
                                                           float* __restrict__ out, int M, int N, int K) {
    extern __shared__ float wt[];                         // [K][NO]
    for (int i = threadIdx.x; i < K * NO; i += blockDim.x) {
        const int k = i / NO, n = i - k * NO;
        wt[i] = (n < N) ? W[(int64_t)n * ldw + k] : 0.0f;
    }
    __syncthreads();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float acc[NO];
#pragma unroll
    for (int n = 0; n < NO; ++n) acc[n] = (n < N) ? bias[n] : 0.0f;
    const uint4* ph = reinterpret_cast<const uint4*>(hs + (int64_t)m * ldh);
    const uint4* pl = reinterpret_cast<const uint4*>(hs + hplane + (int64_t)m * ldh);
    for (int k8 = 0; k8 < K / 8; ++k8) {
        const uint4 a = __ldg(ph + k8), b = __ldg(pl + k8);
        const uint32_t ah[4] = {a.x, a.y, a.z, a.w}, al[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float h0 = bf_lo(ah[t]) + bf_lo(al[t]), h1 = bf_hi(ah[t]) + bf_hi(al[t]);
            const float* w0 = wt + (k8 * 8 + 2 * t) * NO;
#pragma unroll
            for (int n = 0; n < NO; ++n) acc[n] = fmaf(h0, w0[n], acc[n]);
#pragma unroll
            for (int n = 0; n < NO; ++n) acc[n] = fmaf(h1, w0[NO + n], acc[n]);
        }
    }
#pragma unroll
    for (int n = 0; n < NO; ++n)
        if (n < N) out[(int64_t)m * N + n] = acc[n];
}

// One pass over HB_ROWS rows of h (split, width K).  Block = 256 threads = (256-column slab, as 128 column pairs) x 2 row
// groups (a slab narrower than 256 columns packs more row groups: K = 128 -> 64 pairs x 4 groups), so every thread
// is busy; the row groups' partial sums meet in shared memory, then one set of atomics per block.
//   dW[n][k]  += sum_m dY[m][n] h[m][k]                       (head weight gradient)
//   db[n]     += sum_m dY[m][n]                               (head bias gradient; blockIdx.x == 0 only)
//   dZ[m][k]   = (sum_n dY[m][n] W[n][k]) * ELU'(h[m][k])     -> split store
//   dbp[k]    += sum_m dZ[m][k]                               (bias gradient of the layer that produced h)
constexpr int HB_ROWS = 192, HB_THREADS = 256;    // 192 rows: a third of the atomics of 64-row blocks on the same ~1.7 k addresses; static smem 47 KB at NO = 16
template <int NO>
__global__ void __launch_bounds__(HB_THREADS) head_backward_kernel(const float* __restrict__ dY, const uint16_t* __restrict__ hs, int64_t ldh,
                                                                   int64_t hplane, const float* __restrict__ W, int64_t ldw,
                                                                   float* __restrict__ dW, float* __restrict__ db, uint16_t* __restrict__ dzs,
                                                                   int64_t dzplane, float* __restrict__ dbp, int M, int N, int K) {
    __shared__ float dy[HB_ROWS][NO];
    __shared__ float red[HB_THREADS][2 * NO + 2 + 1];        // +1: odd pitch, conflict-free column walks
    const int m0 = blockIdx.y * HB_ROWS, rows = min(HB_ROWS, M - m0);
    for (int i = threadIdx.x; i < HB_ROWS * NO; i += HB_THREADS) {
        const int r = i / NO, n = i - r * NO;
        dy[r][n] = (r < rows && n < N) ? dY[(int64_t)(m0 + r) * N + n] : 0.0f;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < N) {
        float s = 0.0f;
        for (int r = 0; r < rows; ++r) s += dy[r][threadIdx.x];
        atomicAdd(db + threadIdx.x, s);
    }
    const int slab = min(256, K - blockIdx.x * 256);          // columns of this block (even)
    const int pairs = slab >> 1, groups = HB_THREADS / pairs;  // pairs in {4 .. 128}: K % 8 == 0
    const int pi = threadIdx.x % pairs, gi = threadIdx.x / pairs;
    const int k = blockIdx.x * 256 + 2 * pi;
    float a0[NO], a1[NO];
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int n = 0; n < NO; ++n) { a0[n] = 0.0f; a1[n] = 0.0f; }
    if (gi < groups) {
        float w0[NO], w1[NO];
#pragma unroll
        for (int n = 0; n < NO; ++n) {
            w0[n] = (n < N) ? W[(int64_t)n * ldw + k] : 0.0f;
            w1[n] = (n < N) ? W[(int64_t)n * ldw + k + 1] : 0.0f;
        }
        const uint16_t* hp = hs + (int64_t)m0 * ldh + k;
        uint16_t* zp = dzs + (int64_t)m0 * ldh + k;
#pragma unroll 4
        for (int r = gi; r < rows; r += groups) {
            const uint32_t ph = __ldg(reinterpret_cast<const uint32_t*>(hp + (int64_t)r * ldh));
            const uint32_t pl = __ldg(reinterpret_cast<const uint32_t*>(hp + hplane + (int64_t)r * ldh));
            const float h0 = bf_lo(ph) + bf_lo(pl), h1 = bf_hi(ph) + bf_hi(pl);
            float z0 = 0.0f, z1 = 0.0f;
#pragma unroll
            for (int n = 0; n < NO; ++n) {
                const float d = dy[r][n];
                a0[n] = fmaf(d, h0, a0[n]); a1[n] = fmaf(d, h1, a1[n]);
                z0 = fmaf(d, w0[n], z0); z1 = fmaf(d, w1[n], z1);
            }
            z0 *= (h0 > 0.0f) ? 1.0f : (h0 + 1.0f);
            z1 *= (h1 > 0.0f) ? 1.0f : (h1 + 1.0f);
            const uint32_t zh = pack_bf16x2(z0, z1);
            const uint32_t zl = pack_bf16x2(z0 - bf_lo(zh), z1 - bf_hi(zh));
            *reinterpret_cast<uint32_t*>(zp + (int64_t)r * ldh) = zh;
            *reinterpret_cast<uint32_t*>(zp + dzplane + (int64_t)r * ldh) = zl;
            s0 += z0; s1 += z1;
        }
    }
    // meet the row groups' partial sums: thread (pi, gi) parks its 2 NO + 2 values, group 0 adds them up
    float* mine = red[threadIdx.x];
#pragma unroll
    for (int n = 0; n < NO; ++n) { mine[2 * n] = a0[n]; mine[2 * n + 1] = a1[n]; }
    mine[2 * NO] = s0; mine[2 * NO + 1] = s1;
    __syncthreads();
    if (gi == 0) {
#pragma unroll
        for (int n = 0; n < NO; ++n) {
            float t0 = 0.0f, t1 = 0.0f;
            for (int g2 = 0; g2 < groups; ++g2) { t0 += red[g2 * pairs + pi][2 * n]; t1 += red[g2 * pairs + pi][2 * n + 1]; }
            if (n < N) {
                atomicAdd(dW + (int64_t)n * ldw + k, t0);
                atomicAdd(dW + (int64_t)n * ldw + k + 1, t1);
            }
        }
        float t0 = 0.0f, t1 = 0.0f;
        for (int g2 = 0; g2 < groups; ++g2) { t0 += red[g2 * pairs + pi][2 * NO]; t1 += red[g2 * pairs + pi][2 * NO + 1]; }
        atomicAdd(dbp + k, t0);
        atomicAdd(dbp + k + 1, t1);
    }
}

int32_t check_net_split(const HgMlpDesc* net, const HgSplit* X) {
    HG_REQUIRE(net); HG_REQUIRE(X); HG_REQUIRE(X->p);
    if (net->n_layers < 2 || net->n_layers > HG_MAX_LAYERS) return hg_fail(HG_E_ARG, "hg_mlp_*_split: need 2..8 layers");
    const int L = net->n_layers;
    if (net->dims[L] > HEAD_MAX_N) return hg_fail(HG_E_ALIGN, "hg_mlp_*_split: output layer wider than 16");
    for (int l = 1; l < L; ++l)
        if (net->dims[l] & 7) return hg_fail(HG_E_ALIGN, "hg_mlp_*_split: hidden widths must be multiples of 8");
    for (int l = 0; l < L - 1; ++l)
        if ((net->ldw[l] & 7) || (net->w_off[l] & 7)) return hg_fail(HG_E_ALIGN, "hg_mlp_*_split: weight pitch / offset must be multiples of 8");
    if ((X->ld & 7) || (X->plane & 7) || X->ld < net->dims[0]) return hg_fail(HG_E_ALIGN, "hg_mlp_*_split: bad X split layout");
    return 0;
}

HgSplit hidden_slot(uint16_t* base, const HgMlpDesc* net, int l, int64_t M) {   // hidden layer l in 1 .. L-1
    int64_t off = 0;
    for (int j = 1; j < l; ++j) off += 2 * M * net->dims[j];
    HgSplit s;
    s.p = base + off; s.ld = net->dims[l]; s.plane = M * net->dims[l];
    return s;
}

}  // namespace

extern "C" int32_t hg_mlp_forward_split(const HgMlpDesc* net, const float* params, const uint16_t* wsplit, int64_t w_plane,
                                        const HgSplit* X, uint16_t* hidden, float* out, int64_t M, void* stream) {
    if (int32_t rc = check_net_split(net, X)) return rc;
    HG_REQUIRE(params); HG_REQUIRE(wsplit); HG_REQUIRE(hidden); HG_REQUIRE(out);
    if (M <= 0 || M > (1 << 28)) return hg_fail(HG_E_SIZE, "hg_mlp_forward_split: bad M");
    const int L = net->n_layers;
    cudaStream_t st = (cudaStream_t)stream;
    HgSplit in = *X;
    for (int l = 0; l < L - 1; ++l) {
        HgGemmSplit d{};
        d.A = in;
        d.B.p = const_cast<uint16_t*>(wsplit) + net->w_off[l]; d.B.ld = net->ldw[l]; d.B.plane = w_plane;
        d.Cs = hidden_slot(hidden, net, l + 1, M);
        d.bias = params + net->b_off[l];
        d.M = (int32_t)M; d.N = net->dims[l + 1]; d.K = net->dims[l];
        d.epilogue = 2; d.split_k = 1;
        if (int32_t rc = hg_gemm_bf16x3(&d, stream)) return rc;
        in = d.Cs;
    }
    const int K = net->dims[L - 1], N = net->dims[L];
    const float* W = params + net->w_off[L - 1];
    const float* b = params + net->b_off[L - 1];
    const unsigned grid = (unsigned)((M + 127) / 128);
    if ((size_t)K * 16 * sizeof(float) > 48 * 1024) return hg_fail(HG_E_SIZE, "hg_mlp_forward_split: last hidden layer too wide for the head kernel");
#define HEAD_FWD(NO) head_forward_kernel<NO><<<grid, 128, (size_t)K * NO * sizeof(float), st>>>(in.p, in.ld, in.plane, W, net->ldw[L - 1], b, out, (int)M, N, K)
    if (N == 1) HEAD_FWD(1); else if (N <= 4) HEAD_FWD(4); else if (N <= 8) HEAD_FWD(8); else if (N <= 12) HEAD_FWD(12); else HEAD_FWD(16);
#undef HEAD_FWD
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_mlp_forward_split");
}

extern "C" int32_t hg_mlp_backward_split(const HgMlpDesc* net, const float* params, const uint16_t* wsplit, int64_t w_plane,
                                         const HgSplit* X, const uint16_t* hidden, const float* dY, uint16_t* dhidden,
                                         float* grads, int64_t M, void* stream) {
    if (int32_t rc = check_net_split(net, X)) return rc;
    HG_REQUIRE(params); HG_REQUIRE(wsplit); HG_REQUIRE(hidden); HG_REQUIRE(dY); HG_REQUIRE(dhidden); HG_REQUIRE(grads);
    if (M <= 0 || M > (1 << 28)) return hg_fail(HG_E_SIZE, "hg_mlp_backward_split: bad M");
    const int L = net->n_layers;
    cudaStream_t st = (cudaStream_t)stream;
    // zero this net's gradient range (blocks are laid out back to back: w0 b0 w1 b1 ...)
    {
        const int64_t first = net->w_off[0], last = net->b_off[L - 1] + net->dims[L];
        if (last <= first) return hg_fail(HG_E_ARG, "hg_mlp_backward_split: parameter blocks must be laid out in order");
        cudaMemsetAsync(grads + first, 0, sizeof(float) * (size_t)(last - first), st);
    }
    uint16_t* hid = const_cast<uint16_t*>(hidden);
    // ---- output head: dW, db, dZ of hidden layer L-1 (split) and its bias gradient, one pass over h ----
    {
        const int K = net->dims[L - 1], N = net->dims[L];
        const HgSplit h = hidden_slot(hid, net, L - 1, M), dz = hidden_slot(dhidden, net, L - 1, M);
        const float* W = params + net->w_off[L - 1];
        dim3 grid((K + 255) / 256, (unsigned)((M + HB_ROWS - 1) / HB_ROWS));
#define HEAD_BWD(NO) head_backward_kernel<NO><<<grid, HB_THREADS, 0, st>>>(dY, h.p, h.ld, h.plane, W, net->ldw[L - 1], grads + net->w_off[L - 1], \
        grads + net->b_off[L - 1], dz.p, dz.plane, grads + net->b_off[L - 2], (int)M, N, K)
        if (N == 1) HEAD_BWD(1); else if (N <= 4) HEAD_BWD(4); else if (N <= 8) HEAD_BWD(8); else if (N <= 12) HEAD_BWD(12); else HEAD_BWD(16);
#undef HEAD_BWD
        HG_LAUNCHED(1);
        if (int32_t rc = hg_cuda_status("hg_mlp_backward_split (head)")) return rc;
    }
    // ---- hidden layers, last to first: wgrad (split-K, atomics) and dgrad (x ELU', split store, bias gradient) ----
    for (int l = L - 2; l >= 0; --l) {
        const int K = net->dims[l], N = net->dims[l + 1];          // layer l: (M, K) -> (M, N)
        const HgSplit dz = hidden_slot(dhidden, net, l + 1, M);
        const HgSplit in = (l == 0) ? *X : hidden_slot(hid, net, l, M);
        {   // dW[n][k] = sum_m dZ[m][n] in[m][k]
            HgGemmSplit d{};
            d.A = dz; d.B = in;
            d.C = grads + net->w_off[l]; d.ldc = net->ldw[l];
            d.M = N; d.N = K; d.K = (int32_t)M;
            d.a_mn_major = 1; d.b_mn_major = 1; d.epilogue = 4;
            // one wave of work items: 128 x bn tiles on 148 CTAs (split-K launches use the single-CTA form, see hg_gemm_bf16x3)
            const int bn = K >= 256 ? 256 : (K + 63) / 64 * 64;
            const int tiles = ((N + 127) / 128) * ((K + bn - 1) / bn);
            int splits = HG_NUM_SMS / tiles;
            const int64_t cap = (M / 64) / 4;                      // at least 4 k-blocks per work item
            if (splits > cap) splits = (int)cap;
            d.split_k = splits < 1 ? 1 : splits;
            if (int32_t rc = hg_gemm_bf16x3(&d, stream)) return rc;
        }
        if (l > 0) {   // dZ_{l-1} = (dZ_l W_l) * ELU'(h_l);  db_{l-1} = column sums
            HgGemmSplit d{};
            d.A = dz;
            d.B.p = const_cast<uint16_t*>(wsplit) + net->w_off[l]; d.B.ld = net->ldw[l]; d.B.plane = w_plane;
            d.Cs = hidden_slot(dhidden, net, l, M);
            d.Hs = hidden_slot(hid, net, l, M);
            d.colsum = grads + net->b_off[l - 1];
            d.M = (int32_t)M; d.N = K; d.K = N;
            d.a_mn_major = 0; d.b_mn_major = 1; d.epilogue = 3; d.split_k = 1;
            if (int32_t rc = hg_gemm_bf16x3(&d, stream)) return rc;
        }
    }
    return hg_cuda_status("hg_mlp_backward_split");
}
