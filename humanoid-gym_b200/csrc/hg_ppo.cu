// PPO learning-side kernels other than the MLP GEMMs (sm_100a).  All HBM-bound elementwise /
// reduction work; every entry is ONE launch (plus a tiny finalise where a grid-wide value is needed).
//
//   hg_policy_sample      ActorCritic.act + get_actions_log_prob      actor_critic.py:111-120, ppo.py:91-101
//   hg_storage_add        RolloutStorage.add_transitions + bootstrap   rollout_storage.py:87-100, ppo.py:107-108
//   hg_gae                RolloutStorage.compute_returns               rollout_storage.py:122-136
//   hg_minibatch_gather   mini_batch_generator                         rollout_storage.py:146-182
//   hg_ppo_loss_fwd_bwd   PPO.update loss + analytic backward          ppo.py:133-168
//   hg_grad_sqnorm / hg_clip_adam_step   clip_grad_norm_ + Adam        ppo.py:172-173
//   hg_adapt_lr           adaptive-KL schedule                         ppo.py:142-148
#include <stdlib.h>

#include "hg_common.cuh"

namespace {

constexpr float kLogSqrt2Pi = 0.9189385332046727f;      // math.log(math.sqrt(2*math.pi))

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// sample + log-prob
// ---------------------------------------------------------------------------------------------
__global__ void policy_sample_kernel(const float* __restrict__ mean, const float* __restrict__ stdv,
                                     const float* __restrict__ eps, uint64_t seed, uint64_t step,
                                     const uint64_t* __restrict__ step_dev, float* __restrict__ actions, float* __restrict__ logp, float* __restrict__ sigma_out,
                                     int M, int A) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    if (step_dev) step = *step_dev;
    float lp = 0.0f;
    for (int j = 0; j < A; ++j) {
        float mu = mean[(size_t)m * A + j];
        float sg = mu * 0.0f + stdv[j];                        // Normal(mean, mean*0. + std)
        float z;
        if (eps) z = eps[(size_t)m * A + j];
        else {
            HgPhilox r = hg_philox(seed, (uint32_t)m, (uint32_t)step, HG_RNG_SAMPLE | ((uint32_t)(step >> 32) << 8), j);
            z = hg_normal(r.c[0], r.c[1]);
        }
        float a = mu + sg * z;
        float d = a - mu;
        lp += -(d * d) / (2.0f * (sg * sg)) - logf(sg) - kLogSqrt2Pi;
        actions[(size_t)m * A + j] = a;
        sigma_out[(size_t)m * A + j] = sg;
    }
    logp[m] = lp;
}

// ---------------------------------------------------------------------------------------------
// storage add: blockIdx.y selects the tensor, blockIdx.x strides over its elements
// ---------------------------------------------------------------------------------------------
struct AddArgs {
    HgStorage S; HgTransition tr; int t; float gamma; int N;
};
__device__ __forceinline__ void copy_f32(float* dst, const float* src, size_t n) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
        size_t n4 = n >> 2;
        for (size_t i = tid; i < n4; i += nt) reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (size_t i = (n4 << 2) + tid; i < n; i += nt) dst[i] = src[i];
    } else {
        for (size_t i = tid; i < n; i += nt) dst[i] = src[i];
    }
}
__device__ __forceinline__ void copy_slab(float* dst, const float* src, size_t n) {
    if (src != nullptr && src != dst) copy_f32(dst, src, n);
}
// rows of `width` floats from a pitched source into a dense destination
__device__ __forceinline__ void copy_rows(float* dst, const float* src, size_t rows, size_t width, size_t pitch) {
    if (src == nullptr || src == dst) return;
    if (pitch == 0 || pitch == width) { copy_f32(dst, src, rows * width); return; }
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < rows * width; i += nt) {
        size_t r = i / width, c = i - r * width;
        dst[i] = __ldg(src + r * pitch + c);
    }
}
__global__ void storage_add_kernel(AddArgs a) {
    const size_t N = a.N, t = a.t;
    const HgStorage& S = a.S;
    switch (blockIdx.y) {
        case 0: copy_rows(S.observations + t * N * S.num_obs, a.tr.obs, N, S.num_obs, a.tr.obs_pitch); break;
        case 1: if (S.privileged_observations) copy_rows(S.privileged_observations + t * N * S.num_priv, a.tr.priv_obs, N, S.num_priv, a.tr.priv_pitch); break;
        case 2: copy_slab(S.actions + t * N * S.num_actions, a.tr.actions, N * S.num_actions); break;
        case 3: copy_slab(S.mu + t * N * S.num_actions, a.tr.mu, N * S.num_actions); break;
        case 4: copy_slab(S.sigma + t * N * S.num_actions, a.tr.sigma, N * S.num_actions); break;
        default: {
            size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
            for (size_t e = tid; e < N; e += nt) {
                if (a.tr.rewards) {
                    float r = a.tr.rewards[e];
                    if (a.tr.time_outs) {                           // ppo.py:107-108
                        float v = a.tr.values ? a.tr.values[e] : S.values[t * N + e];
                        r += a.gamma * (v * (a.tr.time_outs[e] ? 1.0f : 0.0f));
                    }
                    S.rewards[t * N + e] = r;
                }
                if (a.tr.dones) S.dones[t * N + e] = a.tr.dones[e] ? 1 : 0;
                if (a.tr.values && a.tr.values != S.values + t * N) S.values[t * N + e] = a.tr.values[e];
                if (a.tr.log_prob && a.tr.log_prob != S.actions_log_prob + t * N) S.actions_log_prob[t * N + e] = a.tr.log_prob[e];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GAE: one thread per env walks t = T-1 .. 0 (loads coalesced over envs); block-reduced
// sum / sum-of-squares of the raw advantages in fp64 for the normalisation.
// ---------------------------------------------------------------------------------------------
__global__ void gae_kernel(HgStorage S, const float* __restrict__ last_values, float gamma, float lam,
                           double* __restrict__ stats, int N) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (e < N) {
        float adv = 0.0f;
        float next_v = last_values[e];
        for (int t = S.T - 1; t >= 0; --t) {
            size_t i = (size_t)t * N + e;
            float v = S.values[i];
            float nt = 1.0f - (float)S.dones[i];
            float delta = S.rewards[i] + nt * gamma * next_v - v;
            adv = delta + nt * gamma * lam * adv;
            float ret = adv + v;
            S.returns[i] = ret;
            float a = ret - v;                                  // self.advantages = self.returns - self.values
            S.advantages[i] = a;
            s1 += a; s2 += (double)a * a;
            next_v = v;
        }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    __shared__ double sh[2][8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a += sh[0][k]; b += sh[1][k]; }
        atomicAdd(stats, a); atomicAdd(stats + 1, b);
        if (blockIdx.x == 0) atomicAdd(stats + 2, (double)N * S.T);
    }
}
// GAE as a warp scan over TIME (rollout_storage.py:122-134).  adv_t = d_t + c_t adv_{t+1} with d_t = r_t + nt_t gamma V_{t+1} - V_t
// and c_t = nt_t gamma lambda is a chain of affine maps x -> d + c x, and affine maps compose associatively:
// (c1, d1) o (c2, d2) = (c1 c2, d1 + c1 d2).  A CTA takes 32 envs: the (T, 32) tiles of rewards / values / dones are loaded
// with 128-byte rows (coalesced over envs) into shared memory; then one WARP per env puts time on its lanes -- lane l owns
// the CH = ceil(T / 32) consecutive steps [l CH, (l + 1) CH), folds them into one affine map, a 5-step suffix scan over the
// lanes (shuffles) gives every lane the advantage entering its chunk from the future, and the lane replays its own steps
// with the reference's serial formula.  The 60 dependent steps of the per-env loop become 2 CH + 5.  Rounding differs from
// the serial order only through the scanned carry-in (|c| < 0.9: ~1e-7 relative; the parity bar on returns is 1e-5).
constexpr int GAE_ENVS = 32, GAE_WARPS = 8, GAE_SCAN_MAX_ENVS = 32768;
__global__ void __launch_bounds__(GAE_WARPS * 32) gae_scan_kernel(HgStorage S, const float* __restrict__ last_values, float gamma, float lam,
                                                                  double* __restrict__ stats, int N) {
    // shared tiles [3][32 envs][RS]: rewards -> returns, values, not-terminal -> advantages.  Within an env's row step t sits at
    // (t % CH) * 32 + t / CH, i.e. lane l finds its k-th step at k * 32 + l (conflict-free across the lanes of the scanning
    // warp), and RS = CH * 32 + 1 is odd, so the transposing fill / drain (lanes = envs) is conflict-free too.
    extern __shared__ float sm[];
    const int T = S.T, e0 = blockIdx.x * GAE_ENVS, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int CH = (T + 31) / 32, RS = CH * 32 + 1;
    float* sr = sm;
    float* sv = sm + GAE_ENVS * RS;
    float* sn = sm + 2 * GAE_ENVS * RS;
    auto at = [&](int le, int t) { return le * RS + (t % CH) * 32 + t / CH; };
    for (int i = tid; i < T * GAE_ENVS; i += GAE_WARPS * 32) {
        const int t = i >> 5, le = i & 31;
        const bool ok = e0 + le < N;
        const size_t g = (size_t)t * N + e0 + le;
        const int j = at(le, t);
        sr[j] = ok ? S.rewards[g] : 0.0f;
        sv[j] = ok ? S.values[g] : 0.0f;
        sn[j] = ok ? 1.0f - (float)S.dones[g] : 0.0f;
    }
    __syncthreads();
    double s1 = 0.0, s2 = 0.0;
    for (int le = warp; le < GAE_ENVS && e0 + le < N; le += GAE_WARPS) {
        const float vlast = last_values[e0 + le];
        const int t_lo = lane * CH, t_hi = min(T, t_lo + CH);       // this lane's steps [t_lo, t_hi)
        // fold the chunk (latest step first) into x -> D + C x
        float C = 1.0f, D = 0.0f;
        for (int t = t_hi - 1; t >= t_lo; --t) {
            const int j = at(le, t);
            const float nt = sn[j], v = sv[j];
            const float nv = (t + 1 < T) ? sv[at(le, t + 1)] : vlast;
            const float d = sr[j] + nt * gamma * nv - v;
            const float c = nt * gamma * lam;
            D = d + c * D;                                           // (c, d) o (C, D)
            C = c * C;
        }
        // inclusive suffix scan over the lanes: afterwards (C, D) maps the advantage beyond step T - 1 (= 0) to the advantage
        // at t_lo, so D alone is that advantage
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float C2 = __shfl_down_sync(0xffffffffu, C, o), D2 = __shfl_down_sync(0xffffffffu, D, o);
            if (lane + o < 32) { D = D + C * D2; C = C * C2; }
        }
        float adv = __shfl_down_sync(0xffffffffu, D, 1);              // advantage at the first step of the next lane's chunk
        if (lane == 31) adv = 0.0f;
        for (int t = t_hi - 1; t >= t_lo; --t) {                      // the reference's loop body, rollout_storage.py:124-131
            const int j = at(le, t);
            const float nt = sn[j], v = sv[j];
            const float nv = (t + 1 < T) ? sv[at(le, t + 1)] : vlast;
            const float delta = sr[j] + nt * gamma * nv - v;
            adv = delta + nt * gamma * lam * adv;
            const float ret = adv + v;
            const float a = ret - v;                                  // self.advantages = self.returns - self.values
            sr[j] = ret;
            sn[j] = a;
            s1 += a; s2 += (double)a * a;
        }
    }
    __syncthreads();
    for (int i = tid; i < T * GAE_ENVS; i += GAE_WARPS * 32) {
        const int t = i >> 5, le = i & 31;
        if (e0 + le < N) {
            const size_t g = (size_t)t * N + e0 + le;
            const int j = at(le, t);
            S.returns[g] = sr[j];
            S.advantages[g] = sn[j];
        }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    __shared__ double sh[2][GAE_WARPS];
    if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; }
    __syncthreads();
    if (tid == 0) {
        double a = 0, b = 0;
        for (int k = 0; k < GAE_WARPS; ++k) { a += sh[0][k]; b += sh[1][k]; }
        atomicAdd(stats, a); atomicAdd(stats + 1, b);
        if (blockIdx.x == 0) atomicAdd(stats + 2, (double)N * S.T);
    }
}
__global__ void adv_normalise_kernel(HgStorage S, const double* __restrict__ stats, size_t total) {
    double n = stats[2];
    double mean = stats[0] / n;
    double var = (stats[1] - stats[0] * mean) / (n - 1.0);      // unbiased (torch.std default)
    float m = (float)mean, sd = (float)sqrt(var > 0 ? var : 0.0);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) S.advantages[i] = (S.advantages[i] - m) / (sd + 1e-8f);
}

// ---------------------------------------------------------------------------------------------
// minibatch gather: one warp per row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gather_pack_bf16x2(float x0, float x1) {      // x0 at the lower address
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
// one row of `width` floats -> dense fp32 copy (dst may be NULL) and/or split bf16 planes (sp.p may be NULL);
// a lane owns element pairs (2 lane + 64 i, +1): 128-byte warp stores into each plane
__device__ __forceinline__ void gather_row(const float* __restrict__ src, int width, float* dst, const HgSplit& sp, size_t row, int lane) {
    uint16_t* sh = sp.p ? sp.p + row * sp.ld : nullptr;
    for (int k = 2 * lane; k < width; k += 64) {
        const float x0 = __ldg(src + k), x1 = (k + 1 < width) ? __ldg(src + k + 1) : 0.0f;
        if (dst) {
            dst[k] = x0;
            if (k + 1 < width) dst[k + 1] = x1;
        }
        if (sh) {
            const uint32_t h = gather_pack_bf16x2(x0, x1);
            const uint32_t l = gather_pack_bf16x2(x0 - __uint_as_float(h << 16), x1 - __uint_as_float(h & 0xFFFF0000u));
            *reinterpret_cast<uint32_t*>(sh + k) = h;                     // ld is even (multiple of 8): pair stays inside the row pitch
            *reinterpret_cast<uint32_t*>(sh + sp.plane + k) = l;
        }
    }
}
__global__ void gather_kernel(HgStorage S, const int64_t* __restrict__ idx, HgMiniBatch mb, int B) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= B) return;
    size_t src = (size_t)idx[row];
    gather_row(S.observations + src * S.num_obs, S.num_obs,
               mb.obs ? mb.obs + (size_t)row * (mb.ld_obs ? mb.ld_obs : S.num_obs) : nullptr, mb.obs_split, (size_t)row, lane);
    if (S.privileged_observations)
        gather_row(S.privileged_observations + src * S.num_priv, S.num_priv,
                   mb.priv_obs ? mb.priv_obs + (size_t)row * (mb.ld_priv ? mb.ld_priv : S.num_priv) : nullptr, mb.priv_split, (size_t)row, lane);
    int A = S.num_actions;
    for (int k = lane; k < A; k += 32) {
        mb.actions[(size_t)row * A + k] = S.actions[src * A + k];
        mb.old_mu[(size_t)row * A + k] = S.mu[src * A + k];
        mb.old_sigma[(size_t)row * A + k] = S.sigma[src * A + k];
    }
    if (lane == 0) {
        mb.values[row] = S.values[src];
        mb.advantages[row] = S.advantages[src];
        mb.returns[row] = S.returns[src];
        mb.old_log_prob[row] = S.actions_log_prob[src];
    }
}

// ---------------------------------------------------------------------------------------------
// PPO loss forward + analytic backward, one thread per sample
// ---------------------------------------------------------------------------------------------
constexpr int LOSS_THREADS = 256;
constexpr int MAX_A = 32;

__global__ void __launch_bounds__(LOSS_THREADS) ppo_loss_kernel(HgPpoLossArgs a, int B) {
    const int A = a.num_actions;
    int i = blockIdx.x * LOSS_THREADS + threadIdx.x;
    float sur = 0.0f, vl = 0.0f, kl = 0.0f;
    float dstd[MAX_A];
#pragma unroll
    for (int j = 0; j < MAX_A; ++j) dstd[j] = 0.0f;
    if (i < B) {
        float logp = 0.0f;
        for (int j = 0; j < A; ++j) {
            float mu = a.mean[(size_t)i * A + j], sg = a.std[j], act = a.actions[(size_t)i * A + j];
            float d = act - mu;
            logp += -(d * d) / (2.0f * (sg * sg)) - logf(sg) - kLogSqrt2Pi;
            float os = a.old_sigma[(size_t)i * A + j], om = a.old_mu[(size_t)i * A + j];
            float dm = om - mu;
            kl += logf(sg / os + 1.e-5f) + (os * os + dm * dm) / (2.0f * (sg * sg)) - 0.5f;     // ppo.py:138-139
        }
        float adv = a.advantages[i];
        float ratio = expf(logp - a.old_log_prob[i]);
        float lo = 1.0f - a.clip_param, hi = 1.0f + a.clip_param;
        float rc = fminf(fmaxf(ratio, lo), hi);
        float s1 = -adv * ratio, s2 = -adv * rc;
        sur = fmaxf(s1, s2);
        // d max(s1,s2)/d ratio: ties split evenly (torch maximum); clamp passes grad inside [lo,hi]
        bool inside = (ratio >= lo) && (ratio <= hi);
        float g1 = (s1 > s2) ? 1.0f : (s1 == s2 ? 0.5f : 0.0f);
        float g2 = (s2 > s1) ? 1.0f : (s1 == s2 ? 0.5f : 0.0f);
        float dsur_dratio = -adv * g1 + (inside ? -adv * g2 : 0.0f);
        float dlogp = dsur_dratio * ratio * a.inv_B;           // d loss / d logp_i
        for (int j = 0; j < A; ++j) {
            float mu = a.mean[(size_t)i * A + j], sg = a.std[j], act = a.actions[(size_t)i * A + j];
            float d = act - mu, var = sg * sg;
            a.d_mean[(size_t)i * A + j] = dlogp * (d / var);
            dstd[j] = dlogp * ((d * d) / (var * sg) - 1.0f / sg);
        }
        // value loss, ppo.py:159-166
        float v = a.value[i], tv = a.target_values[i], R = a.returns[i];
        float dv;
        if (a.use_clipped_value_loss) {
            float diff = v - tv;
            float dc = fminf(fmaxf(diff, -a.clip_param), a.clip_param);
            float vc = tv + dc;
            float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
            vl = fmaxf(l1, l2);
            bool in2 = (diff >= -a.clip_param) && (diff <= a.clip_param);
            float h1 = (l1 > l2) ? 1.0f : (l1 == l2 ? 0.5f : 0.0f);
            float h2 = (l2 > l1) ? 1.0f : (l1 == l2 ? 0.5f : 0.0f);
            dv = h1 * 2.0f * (v - R) + (in2 ? h2 * 2.0f * (vc - R) : 0.0f);
        } else {
            vl = (R - v) * (R - v);
            dv = -2.0f * (R - v);
        }
        a.d_value[i] = a.value_loss_coef * dv * a.inv_B;
    }
    // block reductions -> global atomics
    __shared__ float sh[LOSS_THREADS / 32][3 + MAX_A];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    float r0 = warp_sum(sur), r1 = warp_sum(vl), r2 = warp_sum(kl);
    if (l == 0) { sh[w][0] = r0; sh[w][1] = r1; sh[w][2] = r2; }
    for (int j = 0; j < A; ++j) {
        float r = warp_sum(dstd[j]);
        if (l == 0) sh[w][3 + j] = r;
    }
    __syncthreads();
    if (threadIdx.x < 3 + A) {
        float s = 0.0f;
        for (int k = 0; k < LOSS_THREADS / 32; ++k) s += sh[k][threadIdx.x];
        if (threadIdx.x == 0) atomicAdd(a.scalars + 0, s * a.inv_B);
        else if (threadIdx.x == 1) atomicAdd(a.scalars + 1, s * a.inv_B);
        else if (threadIdx.x == 2) atomicAdd(a.scalars + 3, s * a.inv_B);
        else atomicAdd(a.grad_std + (threadIdx.x - 3), s);
    }
}
// entropy term: -entropy_coef * mean_i sum_j (0.5 + 0.5 log 2pi + log sigma_j); identical for all i.
// grad_std must already hold the log-prob part; scalars[2] receives the entropy mean.
__global__ void ppo_entropy_kernel(HgPpoLossArgs a, float batch_fraction) {
    int j = threadIdx.x;
    float ent = 0.0f;
    if (j < a.num_actions) {
        float sg = a.std[j];
        ent = 0.5f + 0.5f * 1.8378770664093453f + logf(sg);
        a.grad_std[j] += -a.entropy_coef * batch_fraction / sg;
    }
    ent = warp_sum(ent);
    if (j == 0) a.scalars[2] = ent * batch_fraction;
}

// ---------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam
// ---------------------------------------------------------------------------------------------
__global__ void sqnorm_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = g[i];
        s += v * v;
    }
    s = warp_sum(s);
    __shared__ double sh[8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
        atomicAdd(out, t);
    }
}

struct AdamArgs {
    float* p; const float* g; float* m; float* v; const double* sqnorm; const double* lr; int32_t* step;
    float max_norm, beta1, beta2, eps, grad_scale; int64_t n;
};
__global__ void clip_adam_kernel(AdamArgs a) {
    // clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
    float total_norm = (float)(sqrt(*a.sqnorm) * (double)a.grad_scale);
    float coef = fminf(a.max_norm / (total_norm + 1e-6f), 1.0f) * a.grad_scale;
    const int t = *a.step + 1;                                   // this optimizer step (1-based)
    const double bc1 = 1.0 - pow((double)a.beta1, (double)t);
    const double bc2 = 1.0 - pow((double)a.beta2, (double)t);
    const float step_size = (float)(*a.lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - (double)a.beta1), w2 = (float)(1.0 - (double)a.beta2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        float g = a.g[i] * coef;
        float m = a.m[i], v = a.v[i];
        m = m + w1 * (g - m);                                    // exp_avg.lerp_(grad, 1-beta1)
        v = v * a.beta2 + w2 * (g * g);                          // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
        float denom = sqrtf(v) / bc2_sqrt + a.eps;
        a.p[i] = a.p[i] - step_size * (m / denom);               // param.addcdiv_(exp_avg, denom, -step_size)
        a.m[i] = m; a.v[i] = v;
    }
}
// one thread after the update: Adam step count, sqnorm re-zeroed, and (optionally) this minibatch's loss statistics added to
// the running sums PPO.update reports (ppo.py:175-176: mean_value_loss += ..., mean_surrogate_loss += ...)
__global__ void adam_finish_kernel(int32_t* step, double* sqnorm, const float* stats, float* stats_sum, int n_stats) {
    *step += 1;
    *sqnorm = 0.0;
    if (stats_sum)
        for (int k = 0; k < n_stats; ++k) stats_sum[k] += stats[k];
}

// ---------------------------------------------------------------------------------------------
// minibatch permutation (rollout_storage.py:155: torch.randperm) without a sort: a keyed bijection of [0, 2^b) --
// an alternating unbalanced Feistel network, 8 half-rounds with a murmur3-finalizer round function -- walked along
// its cycle until it lands inside [0, n) ("cycle walking"); 2^b < 2n, so fewer than 2 evaluations on average.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__global__ void randperm_kernel(int64_t n, int bits_lo, int bits_hi, uint64_t seed, uint64_t counter, int64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t mlo = (bits_lo >= 32) ? 0xFFFFFFFFu : ((1u << bits_lo) - 1u);
    const uint32_t mhi = (bits_hi >= 32) ? 0xFFFFFFFFu : ((1u << bits_hi) - 1u);
    const HgPhilox kx = hg_philox(seed, (uint32_t)counter, (uint32_t)(counter >> 32), 0x50455246u, 0);   // round keys
    const HgPhilox ky = hg_philox(seed, (uint32_t)counter, (uint32_t)(counter >> 32), 0x50455246u, 1);
    const uint32_t key[8] = {kx.c[0], kx.c[1], kx.c[2], kx.c[3], ky.c[0], ky.c[1], ky.c[2], ky.c[3]};
    uint64_t x = (uint64_t)i;
    do {
        uint32_t lo = (uint32_t)x & mlo, hi = (uint32_t)(x >> bits_lo) & mhi;
#pragma unroll
        for (int r = 0; r < 8; r += 2) {
            hi ^= fmix32(lo ^ key[r]) & mhi;
            lo ^= fmix32(hi ^ key[r + 1]) & mlo;
        }
        x = ((uint64_t)hi << bits_lo) | lo;
    } while (x >= (uint64_t)n);
    out[i] = (int64_t)x;
}

// OnPolicyRunner.learn's per-step bookkeeping (on_policy_runner.py:140-154) for one env step: running reward / length of the
// current episode per env; where the env finished, the totals go to row t of (T, N) result slabs (NaN elsewhere) and the
// running values restart; the step's extras["episode"] means are filed into row t of a (T, n_infos) slab.
__global__ void episode_book_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones, float* __restrict__ cur_rew,
                                    float* __restrict__ cur_len, float* __restrict__ done_rew, float* __restrict__ done_len,
                                    const float* __restrict__ infos_in, float* __restrict__ infos_out, int n_infos, int N) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_infos && infos_in) infos_out[e] = infos_in[e];
    if (e >= N) return;
    const float r = cur_rew[e] + rewards[e];                    // cur_reward_sum += rewards
    const float l = cur_len[e] + 1.0f;                          // cur_episode_length += 1
    const bool d = dones[e] != 0;
    const float nanv = __int_as_float(0x7fc00000);
    done_rew[e] = d ? r : nanv;
    done_len[e] = d ? l : nanv;
    cur_rew[e] = d ? 0.0f : r;
    cur_len[e] = d ? 0.0f : l;
}

__global__ void adapt_lr_kernel(const float* kl_mean, double desired_kl, double* lr) {
    float kl = *kl_mean;                                         // ppo.py:142-145 (fp32 tensor vs python scalar)
    double v = *lr;
    if (kl > (float)(desired_kl * 2.0)) v = fmax(1e-5, v / 1.5);
    else if (kl < (float)(desired_kl / 2.0) && kl > 0.0f) v = fmin(1e-2, v * 1.5);
    *lr = v;
}

}  // namespace

extern "C" int32_t hg_policy_sample(const float* mean, const float* std, const float* eps, uint64_t seed, uint64_t step,
                                    const uint64_t* step_dev, float* actions, float* log_prob, float* sigma_out, int64_t M,
                                    int32_t A, void* stream) {
    HG_REQUIRE(mean); HG_REQUIRE(std); HG_REQUIRE(actions); HG_REQUIRE(log_prob); HG_REQUIRE(sigma_out);
    if (M <= 0 || A <= 0 || A > MAX_A) return hg_fail(HG_E_SIZE, "hg_policy_sample: bad M/A");
    policy_sample_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(mean, std, eps, seed, step, step_dev, actions,
                                                                                      log_prob, sigma_out, (int)M, A);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_policy_sample");
}

extern "C" int32_t hg_storage_add(const HgStorage* S, const HgTransition* tr, int32_t t, float gamma, int64_t N, void* stream) {
    HG_REQUIRE(S); HG_REQUIRE(tr);
    HG_REQUIRE(S->observations); HG_REQUIRE(S->actions); HG_REQUIRE(S->rewards); HG_REQUIRE(S->dones); HG_REQUIRE(S->values);
    HG_REQUIRE(S->actions_log_prob); HG_REQUIRE(S->mu); HG_REQUIRE(S->sigma);
    if (N <= 0) return hg_fail(HG_E_SIZE, "hg_storage_add: bad N");
    if (t < 0 || t >= S->T) return hg_fail(HG_E_STATE, "Rollout buffer overflow");   // rollout_storage.py:88-89
    AddArgs a{*S, *tr, t, gamma, (int)N};
    unsigned gx = (unsigned)((N * S->num_obs / 4 + 255) / 256);
    if (gx > 2 * HG_NUM_SMS) gx = 2 * HG_NUM_SMS;
    if (gx < 1) gx = 1;
    storage_add_kernel<<<dim3(gx, 6), 256, 0, (cudaStream_t)stream>>>(a);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_storage_add");
}

extern "C" int32_t hg_adv_normalise(const HgStorage* S, const double* stats, int64_t N, void* stream) {
    HG_REQUIRE(S); HG_REQUIRE(stats); HG_REQUIRE(S->advantages);
    size_t total = (size_t)N * S->T;
    adv_normalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*S, stats, total);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_adv_normalise");
}

static int g_gae_scan = -1;      // -1: take HG_GAE from the environment on first use; 0 serial, 1 scan, 2 auto (by N)
extern "C" int32_t hg_set_gae_mode(int32_t scan) {
    const int prev = g_gae_scan;
    g_gae_scan = scan < 0 ? -1 : (scan > 2 ? 2 : scan);
    return prev;
}

extern "C" int32_t hg_gae(const HgStorage* S, const float* last_values, float gamma, float lam, double* stats,
                          int32_t normalise, int64_t N, void* stream) {
    HG_REQUIRE(S); HG_REQUIRE(last_values); HG_REQUIRE(stats);
    HG_REQUIRE(S->rewards); HG_REQUIRE(S->values); HG_REQUIRE(S->dones); HG_REQUIRE(S->returns); HG_REQUIRE(S->advantages);
    if (N <= 0 || S->T <= 0) return hg_fail(HG_E_SIZE, "hg_gae: bad N/T");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(stats, 0, 4 * sizeof(double), st);
    // HG_GAE=scan | serial | auto (default).  auto: the warp scan over time for N <= GAE_SCAN_MAX_ENVS, where a thread-per-env walk
    // leaves SMs idle behind 60 dependent steps (compute_returns at T = 60, scan vs serial: 20.5 vs 45.0 us at N = 4096, 28.5 vs
    // 49.1 us at 16384), the serial walk above that (65.4 vs 62.1 us at 65536: enough envs to fill the machine, no transposition).
    if (g_gae_scan == -1) {
        const char* v = getenv("HG_GAE");
        g_gae_scan = (v && !strcmp(v, "serial")) ? 0 : ((v && !strcmp(v, "scan")) ? 1 : 2);
    }
    const int scan = g_gae_scan == 2 ? (N <= GAE_SCAN_MAX_ENVS) : g_gae_scan;
    const size_t smem = (size_t)3 * GAE_ENVS * (((S->T + 31) / 32) * 32 + 1) * sizeof(float);
    if (scan && smem <= 200 * 1024) {
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaFuncSetAttribute(gae_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_set[dev] = true;
        }
        gae_scan_kernel<<<(unsigned)((N + GAE_ENVS - 1) / GAE_ENVS), GAE_WARPS * 32, smem, st>>>(*S, last_values, gamma, lam, stats, (int)N);
    } else {
        gae_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(*S, last_values, gamma, lam, stats, (int)N);
    }
    HG_LAUNCHED(1);
    if (normalise) return hg_adv_normalise(S, stats, N, stream);
    return hg_cuda_status("hg_gae");
}

extern "C" int32_t hg_minibatch_gather(const HgStorage* S, const int64_t* idx, const HgMiniBatch* mb, int64_t B, void* stream) {
    HG_REQUIRE(S); HG_REQUIRE(idx); HG_REQUIRE(mb);
    HG_REQUIRE(mb->actions); HG_REQUIRE(mb->values); HG_REQUIRE(mb->advantages); HG_REQUIRE(mb->returns);
    HG_REQUIRE(mb->old_log_prob); HG_REQUIRE(mb->old_mu); HG_REQUIRE(mb->old_sigma);
    if (!mb->obs && !mb->obs_split.p) return hg_fail(HG_E_NULL, "hg_minibatch_gather: neither obs nor obs_split given");
    if (S->privileged_observations && !mb->priv_obs && !mb->priv_split.p) return hg_fail(HG_E_NULL, "hg_minibatch_gather: neither priv_obs nor priv_split given");
    if ((mb->obs_split.p && ((mb->obs_split.ld & 7) || mb->obs_split.ld < S->num_obs + (S->num_obs & 1))) ||
        (mb->priv_split.p && ((mb->priv_split.ld & 7) || mb->priv_split.ld < S->num_priv + (S->num_priv & 1))))
        return hg_fail(HG_E_ALIGN, "hg_minibatch_gather: split row pitch must be a multiple of 8 and cover the row");
    if (B <= 0) return hg_fail(HG_E_SIZE, "hg_minibatch_gather: bad B");
    gather_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(*S, idx, *mb, (int)B);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_minibatch_gather");
}

extern "C" int32_t hg_ppo_loss_fwd_bwd(const HgPpoLossArgs* a, int64_t B, void* stream) {
    HG_REQUIRE(a);
    HG_REQUIRE(a->mean); HG_REQUIRE(a->value); HG_REQUIRE(a->std); HG_REQUIRE(a->actions); HG_REQUIRE(a->target_values);
    HG_REQUIRE(a->advantages); HG_REQUIRE(a->returns); HG_REQUIRE(a->old_log_prob); HG_REQUIRE(a->old_mu); HG_REQUIRE(a->old_sigma);
    HG_REQUIRE(a->d_mean); HG_REQUIRE(a->d_value); HG_REQUIRE(a->grad_std); HG_REQUIRE(a->scalars);
    if (B <= 0 || a->num_actions <= 0 || a->num_actions > MAX_A) return hg_fail(HG_E_SIZE, "hg_ppo_loss_fwd_bwd: bad B/A");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(a->scalars, 0, 8 * sizeof(float), st);
    cudaMemsetAsync(a->grad_std, 0, a->num_actions * sizeof(float), st);
    ppo_loss_kernel<<<(unsigned)((B + LOSS_THREADS - 1) / LOSS_THREADS), LOSS_THREADS, 0, st>>>(*a, (int)B);
    ppo_entropy_kernel<<<1, 32, 0, st>>>(*a, (float)((double)B * (double)a->inv_B));
    HG_LAUNCHED(2);
    return hg_cuda_status("hg_ppo_loss_fwd_bwd");
}

extern "C" int32_t hg_grad_sqnorm(const float* grads, int64_t n, double* sqnorm_out, void* stream) {
    HG_REQUIRE(grads); HG_REQUIRE(sqnorm_out);
    if (n <= 0) return hg_fail(HG_E_SIZE, "hg_grad_sqnorm: bad n");
    unsigned grid = (unsigned)((n + 1023) / 1024);
    if (grid > 4 * HG_NUM_SMS) grid = 4 * HG_NUM_SMS;
    sqnorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grads, n, sqnorm_out);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_grad_sqnorm");
}

extern "C" int32_t hg_clip_adam_step_stats(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                           double* sqnorm, float max_grad_norm, const double* lr_dev, int32_t* step_dev,
                                           float beta1, float beta2, float eps, float grad_scale, int64_t n,
                                           const float* stats, float* stats_sum, int32_t n_stats, void* stream) {
    if (stats_sum && (!stats || n_stats <= 0 || n_stats > 64)) return hg_fail(HG_E_ARG, "hg_clip_adam_step_stats: stats_sum without stats");
    HG_REQUIRE(params); HG_REQUIRE(grads); HG_REQUIRE(exp_avg); HG_REQUIRE(exp_avg_sq); HG_REQUIRE(sqnorm);
    HG_REQUIRE(lr_dev); HG_REQUIRE(step_dev);
    if (n <= 0) return hg_fail(HG_E_SIZE, "hg_clip_adam_step: bad n");
    cudaStream_t st = (cudaStream_t)stream;
    AdamArgs a{params, grads, exp_avg, exp_avg_sq, sqnorm, lr_dev, step_dev, max_grad_norm, beta1, beta2, eps, grad_scale, n};
    unsigned grid = (unsigned)((n + 255) / 256);
    if (grid > 8 * HG_NUM_SMS) grid = 8 * HG_NUM_SMS;
    clip_adam_kernel<<<grid, 256, 0, st>>>(a);
    adam_finish_kernel<<<1, 1, 0, st>>>(step_dev, sqnorm, stats, stats_sum, n_stats);
    HG_LAUNCHED(2);
    return hg_cuda_status("hg_clip_adam_step");
}

extern "C" int32_t hg_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                     double* sqnorm, float max_grad_norm, const double* lr_dev, int32_t* step_dev,
                                     float beta1, float beta2, float eps, float grad_scale, int64_t n, void* stream) {
    return hg_clip_adam_step_stats(params, grads, exp_avg, exp_avg_sq, sqnorm, max_grad_norm, lr_dev, step_dev, beta1, beta2, eps,
                                   grad_scale, n, nullptr, nullptr, 0, stream);
}

extern "C" int32_t hg_randperm(int64_t n, uint64_t seed, uint64_t counter, int64_t* out, void* stream) {
    HG_REQUIRE(out);
    if (n <= 0 || n > ((int64_t)1 << 40)) return hg_fail(HG_E_SIZE, "hg_randperm: bad n");
    int bits = 1;
    while (((int64_t)1 << bits) < n) ++bits;
    if (bits < 2) bits = 2;
    const int bits_lo = bits / 2, bits_hi = bits - bits_lo;
    randperm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, bits_lo, bits_hi, seed, counter, out);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_randperm");
}

extern "C" int32_t hg_episode_book_step(const float* rewards, const uint8_t* dones, float* cur_reward_sum, float* cur_episode_length,
                                        float* done_rew_t, float* done_len_t, const float* infos_in, float* infos_out_t,
                                        int32_t n_infos, int64_t N, void* stream) {
    HG_REQUIRE(rewards); HG_REQUIRE(dones); HG_REQUIRE(cur_reward_sum); HG_REQUIRE(cur_episode_length); HG_REQUIRE(done_rew_t); HG_REQUIRE(done_len_t);
    if (N <= 0 || n_infos < 0 || (n_infos > 0 && infos_in && !infos_out_t)) return hg_fail(HG_E_SIZE, "hg_episode_book_step: bad N / infos");
    const int64_t work = N > n_infos ? N : n_infos;
    episode_book_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, dones, cur_reward_sum, cur_episode_length,
                                                                                         done_rew_t, done_len_t, infos_in, infos_out_t, n_infos, (int)N);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_episode_book_step");
}

extern "C" int32_t hg_adapt_lr(const float* kl_mean_dev, double desired_kl, double* lr_dev, void* stream) {
    HG_REQUIRE(kl_mean_dev); HG_REQUIRE(lr_dev);
    adapt_lr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(kl_mean_dev, desired_kl, lr_dev);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_adapt_lr");
}
