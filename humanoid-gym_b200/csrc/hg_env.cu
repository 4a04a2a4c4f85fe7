// Environment-side kernels of the humanoid_ppo hot path (sm_100a).
//
//   hg_env_pre_physics     XBotLFreeEnv.step prologue        humanoid_env.py:189-197, legged_robot.py:90-91
//   hg_env_compute_torques LeggedRobot._compute_torques      legged_robot.py:340-356
//   hg_env_post_physics    LeggedRobot.post_physics_step and everything it calls, ONE launch:
//                          legged_robot.py:119-235,304-336,359-397 + humanoid_env.py:83-142,200-540
//
// Data layout: the reference's own env-major row-major tensors (Isaac Gym tensor API), untouched.
// A CTA owns 32 consecutive envs (one warp's worth).  Rows of consecutive envs are contiguous, so every
// per-CTA tile of per-env state (root 13, dof 24, actions 12, contacts 39, ... floats per env) is ONE
// contiguous 16-byte-aligned range: all 8 warps stage those tiles into shared memory with coalesced
// cp.async copies.  A CTA is small (4 warps, ~27 KB of shared memory: tiles that are dead by the time the new
// observation frames are assembled are aliased with them) so that 7 CTAs -- 7 compute warps -- are resident per SM;
// the kernel's critical path is the ~4k dependent instructions of the per-env program, and only many concurrent
// compute warps hide it.  Then the CTA splits by role:
//   * warp 0, one lane per env, evaluates the branchy / transcendental-heavy step entirely out of shared
//     memory (no dependent global loads on its critical path) and publishes the per-env reset flags early;
//   * warps 1-3 stream the observation histories -- 87 % of the kernel's bytes: obs_out[e][0:658] =
//     obs_in[e][47:705], priv likewise -- straight through registers with many independent 128-byte
//     requests in flight; input and output histories are distinct (ping-pong) buffers, so the shift has no
//     in-place hazard and needs no staging.
// The newest frame (47 / 73 floats per env), the observation noise and the +-18 clip are applied by all
// warps after the compute warp finishes.
//
// Compiled with -fmad=false: the reference is a chain of separate fp32 torch ops, so every multiply and add
// rounds on its own; op order below follows the cited lines.
#include "hg_common.cuh"

#define HG_ENVS_PER_CTA 32
#define HG_ENV_THREADS 128
#define HG_MAX_BODIES 16

namespace {

constexpr int E = HG_ENVS_PER_CTA;
constexpr int NCW = 4;                              // compute warps ("roles") per 32-env tile, one lane per env each
constexpr int OBS_W = HG_OBS1 * HG_OBS_FRAMES;      // 705
constexpr int PRIV_W = HG_PRIV1 * HG_PRIV_FRAMES;   // 219
constexpr int OBS_KEEP = OBS_W - HG_OBS1;           // 658 floats survive the shift
constexpr int PRIV_KEEP = PRIV_W - HG_PRIV1;        // 146
constexpr float kTwoPi = 6.283185307179586f;        // float32(2*np.pi)
constexpr float kPi = 3.141592653589793f;           // float32(np.pi)

// The env constants (HgEnvParams, ~0.8 KB) travel BY VALUE as a __grid_constant__ kernel parameter: same constant-bank operand
// access as a __constant__ global, but per launch -- nothing process-global that a second env instance (or another device)
// could overwrite under a captured CUDA graph, and no upload before the launch.

struct V3 { float x, y, z; };

// isaacgym.torch_utils.quat_rotate_inverse (xyzw):  v(2w^2-1) - 2w(u x v) + 2u(u.v)
__device__ __forceinline__ V3 quat_rotate_inverse(const float* q, V3 v) {
    float w = q[3], ux = q[0], uy = q[1], uz = q[2];
    float s = 2.0f * (w * w) - 1.0f;
    V3 a = {v.x * s, v.y * s, v.z * s};
    V3 c = {uy * v.z - uz * v.y, uz * v.x - ux * v.z, ux * v.y - uy * v.x};
    V3 b = {c.x * w * 2.0f, c.y * w * 2.0f, c.z * w * 2.0f};
    float d = ux * v.x + uy * v.y + uz * v.z;
    V3 e = {ux * d * 2.0f, uy * d * 2.0f, uz * d * 2.0f};
    return {a.x - b.x + e.x, a.y - b.y + e.y, a.z - b.z + e.z};
}

// torch.remainder for a positive divisor.  fmod is exact, and for |a| < m it is the identity: that covers
// every angle on this path (atan2 / asin outputs, heading errors); the general case is kept out of line.
__device__ __noinline__ float fmod_general(float a, float m) { return fmodf(a, m); }
__device__ __forceinline__ float remainder_pos(float a, float m) {
    float r = (fabsf(a) < m) ? a : fmod_general(a, m);
    if (r != 0.0f && r < 0.0f) r += m;
    return r;
}
__device__ __noinline__ float atan2_call(float y, float x) { return atan2f(y, x); }

// get_euler_xyz + fold of (pi, 2pi) to negative angles (legged_robot.py:50-55)
__device__ __noinline__ V3 euler_xyz_wrapped(const float* q) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float roll = atan2_call(2.0f * (w * x + y * z), w * w - x * x - y * y + z * z);
    float sinp = 2.0f * (w * y - z * x);
    float pitch = (fabsf(sinp) >= 1.0f) ? copysignf(1.5707963267948966f, sinp) : asinf(sinp);
    float yaw = atan2_call(2.0f * (w * z + x * y), w * w + x * x - y * y - z * z);
    V3 e = {remainder_pos(roll, kTwoPi), remainder_pos(pitch, kTwoPi), remainder_pos(yaw, kTwoPi)};
    if (e.x > kPi) e.x -= kTwoPi;
    if (e.y > kPi) e.y -= kTwoPi;
    if (e.z > kPi) e.z -= kTwoPi;
    return e;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// (exp(-100|clamp(d-lo,-.5,0)|) + exp(-100|clamp(d-hi,0,.5)|))/2, humanoid_env.py:282-305
__device__ __forceinline__ float two_point_distance_reward(float ax, float ay, float bx, float by, float lo, float hi) {
    float dx = ax - bx, dy = ay - by;
    float d = sqrtf(dx * dx + dy * dy);
    float dmin = clampf(d - lo, -0.5f, 0.0f);
    float dmax = clampf(d - hi, 0.0f, 0.5f);
    return (expf(-fabsf(dmin) * 100.0f) + expf(-fabsf(dmax) * 100.0f)) / 2.0f;
}

// cooperative (whole CTA) ASYNCHRONOUS global -> shared copy of a contiguous float range (cp.async / LDGSTS):
// no register round trip, so the staging loop issues every tile's requests back to back and pays the
// memory latency once (cp_async_wait_all + __syncthreads) instead of once per tile.
__device__ __forceinline__ void cp_async4(float* sdst, const float* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(float* sdst, const float* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// the calling thread's arrival on `bar` is triggered when all of its earlier cp.async copies have landed
__device__ __forceinline__ void mbar_arrive_on_cp_async(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(sa), "r"(parity) : "memory");
}
__device__ __forceinline__ void tile_load(float* s, const float* g, int n, int t, int nt) {
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        int n4 = n >> 2;
        for (int i = t; i < n4; i += nt) cp_async16(s + 4 * i, g + 4 * i);
        for (int i = (n4 << 2) + t; i < n; i += nt) cp_async4(s + i, g + i);
    } else {
        for (int i = t; i < n; i += nt) cp_async4(s + i, g + i);
    }
}

// every array length is a multiple of 4 floats so that each member stays 16-byte aligned.
// Aliasing: the tiles in `in_a` / `rg` are last read by the reward section; after a __syncwarp the observation
// section reuses their storage for the new privileged / actor frames.
constexpr int HG_CF_SLOTS = 2 + 2 * HG_MAX_CONTACT_BODIES;          // feet, termination bodies, penalised bodies
struct InA {
    float lact[E * 12];
    float llact[E * 12];
    float ldv[E * 12];
    float tau[E * 12];
    float ref[E * 12];
    float lrv[E * 6];
    float cmd[E * 4];
    float fat[E * 2], fh[E * 2], lfz[E * 2];
};
struct __align__(16) EnvSmem {
    float root[E * 13];
    float dof[E * 24];
    float act[E * 12];
    union {
        InA in_a;                                                   // 2432 floats
        float newpriv[E * HG_PRIV1];                                // 2336 floats
    };
    union {
        float rg[E * 4 * 13];                                       // feet L/R, knee L/R rows of rigid_state
        float newobs[E * HG_OBS1];
    };
    float cf[E * HG_CF_SLOTS * 3];                                  // contact forces of the bodies the step reads
    float sums[HG_NUM_REWARDS * E];                                 // episode sums tile, (22, 32)
    float rpf[E * 3], rpt[E * 3], org[E * 3];
    float fric[E], mass[E];
    float acc[HG_NUM_REWARDS + 2];
    long long ep[E];
    unsigned long long mbar;
    int cnt;
    int is_last;
    int next_row[2];                                                // history-shift row queues (actor / privileged)
    int next_noise;                                                 // observation-noise work queue (chunks of 32 channel pairs)
    float z[E * HG_OBS1];                                           // N(0,1) draws of the new actor frame
    struct {                                                        // exchange block of the four compute roles, [.][env]
        float blv[3][E], bav[3][E], pg[3][E], eul[3][E], s_r[E], c_r[E], cmd[4][E], fat[2][E], rk[HG_NUM_REWARDS][E];
        unsigned char reset_any[E];
    } x;
    unsigned char lc[E * 2];
    unsigned char reset_in[E];
    unsigned char reset[E];
    unsigned char root_dirty[E];
};
static_assert(sizeof(InA) >= sizeof(float) * E * HG_PRIV1, "aliased region too small for the privileged frame");
static_assert(sizeof(float) * E * 4 * 13 >= sizeof(float) * E * HG_OBS1, "aliased region too small for the actor frame");

#define FIELD(f) (int)(offsetof(HgEnvBuffers, f) / sizeof(void*))
#define SMEMF(f) (int)(offsetof(EnvSmem, f) / sizeof(float))
constexpr int kNumTiles = 18;
__constant__ int kTileField[kNumTiles] = {
    FIELD(root_states), FIELD(dof_state), FIELD(actions), FIELD(last_actions), FIELD(last_last_actions), FIELD(last_dof_vel),
    FIELD(torques), FIELD(ref_dof_pos), FIELD(last_root_vel), FIELD(commands), FIELD(feet_air_time),
    FIELD(feet_height), FIELD(last_feet_z), FIELD(rand_push_force), FIELD(rand_push_torque), FIELD(env_origins),
    FIELD(env_frictions), FIELD(body_mass)};
__constant__ int kTileSmem[kNumTiles] = {
    SMEMF(root), SMEMF(dof), SMEMF(act), SMEMF(in_a.lact), SMEMF(in_a.llact), SMEMF(in_a.ldv), SMEMF(in_a.tau), SMEMF(in_a.ref),
    SMEMF(in_a.lrv), SMEMF(in_a.cmd), SMEMF(in_a.fat), SMEMF(in_a.fh), SMEMF(in_a.lfz), SMEMF(rpf), SMEMF(rpt), SMEMF(org),
    SMEMF(fric), SMEMF(mass)};
__constant__ int kTileWidth[kNumTiles] = {13, 24, 12, 12, 12, 12, 12, 12, 6, 4, 2, 2, 2, 3, 3, 3, 1, 1};

__device__ __noinline__ HgPhilox philox_call(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return hg_philox(seed, c0, c1, c2, c3);
}

__device__ __forceinline__ float draw_u(const float* inj, int64_t idx, uint64_t seed, uint64_t step, uint32_t env,
                                        uint32_t stream, uint32_t k) {
    if (inj) return inj[idx];
    HgPhilox r = philox_call(seed, env, (uint32_t)step, stream | ((uint32_t)(step >> 32) << 8), k);
    return hg_u01(r.c[0]);
}

// legged_robot.py:322-336  (torch_rand_float(lo,hi) = (hi-lo)*u + lo; spans are formed in double on the host)
__device__ __forceinline__ void resample_commands(const HgEnvParams& cP, float* cmd, float u0, float u1, float u2) {
    cmd[0] = cP.cmd_x_span * u0 + cP.cmd_x_lo;
    cmd[1] = cP.cmd_y_span * u1 + cP.cmd_y_lo;
    cmd[3] = cP.cmd_heading_span * u2 + cP.cmd_heading_lo;
    float nrm = sqrtf(cmd[0] * cmd[0] + cmd[1] * cmd[1]);
    float keep = nrm > 0.2f ? 1.0f : 0.0f;
    cmd[0] *= keep;
    cmd[1] *= keep;
}


// History shift of one CTA tile: out[r][0:KEEP] = in[r][FRAME:FRAME+KEEP].  Warps pull batches of ROWS rows off a shared-memory
// queue (self-balancing: the compute warps join late), each batch = ROWS * ceil(KEEP/32) independent 128-byte requests in
// flight per warp, straight through registers (input and output histories are distinct ping-pong buffers: no hazard).
template <int FRAME, int KEEP, int ROWS>
__device__ __forceinline__ void stream_history(float* __restrict__ out, const float* __restrict__ in, int* next_row, int nE, int lane,
                                               int64_t pitch) {
    constexpr int CH = (KEEP + 31) / 32;            // 21 (obs) / 5 (priv)
#pragma unroll 1
    for (;;) {
        int r0 = 0;
        if (lane == 0) r0 = atomicAdd(next_row, ROWS);
        r0 = __shfl_sync(0xffffffffu, r0, 0);
        if (r0 >= nE) break;
        float v[ROWS][CH];
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const int r = r0 + q;
            const bool live = (r < nE);
            const float* src = in + (size_t)r * pitch + FRAME;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                int k = c * 32 + lane;
                v[q][c] = (live && k < KEEP) ? __ldcs(src + k) : 0.0f;      // streaming: read once
            }
        }
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const int r = r0 + q;
            if (r < nE) {
                float* dst = out + (size_t)r * pitch;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    int k = c * 32 + lane;
                    if (k < KEEP) dst[k] = v[q][c];
                }
            }
        }
    }
}

// T = threads per CTA: 4 compute warps (4 lanes per env, 8 envs per warp) that join the history shift when done, plus T/32 - 4
// pure streaming warps: 128 (no extra warps) when the grid is several waves deep, 512 when there is at most one tile per SM
#ifndef HIST_OBS_ROWS
#define HIST_OBS_ROWS 2
#endif
#ifndef HIST_CTAS_128
#define HIST_CTAS_128 4
#endif
// debug timeline (hg_env_set_trace): thread 0 of every CTA stamps %globaltimer at the phase boundaries, [grid][12] int64
#define ETRACE(slot_) do { if (trace && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); trace[(size_t)blockIdx.x * 12 + (slot_)] = t_; } } while (0)
template <int T>
__global__ void __launch_bounds__(T, (T == 128 ? HIST_CTAS_128 : (T == 256 ? 2 : 1)))
post_physics_kernel(const __grid_constant__ HgEnvParams cP, HgEnvBuffers B, HgEnvNoise Z, uint32_t phases, int64_t common_step, int N,
                    long long* trace, int l2_prefetch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EnvSmem& S = *reinterpret_cast<EnvSmem*>(smem_raw);
    auto& X = S.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_tiles = (N + E - 1) / E;
    const bool do_obs = phases & HG_PHASE_OBS, do_reset = phases & HG_PHASE_RESET, do_last = phases & HG_PHASE_LAST;
    const int nb = cP.num_bodies;
    const int64_t opitch = B.obs_pitch ? B.obs_pitch : OBS_W, ppitch = B.priv_pitch ? B.priv_pitch : PRIV_W;
    if (Z.use_device_counters) {   // CUDA-graph friendly: counters live in scratch[4..7], bumped by the last CTA
        common_step = *reinterpret_cast<const volatile int64_t*>(B.scratch + 4) + ((phases & HG_PHASE_COUNTERS) ? 1 : 0);
        Z.step = *reinterpret_cast<const volatile uint64_t*>(B.scratch + 6);
    }

    // ---- 1. stage every per-env input: all 8 warps ISSUE the asynchronous copies, nobody blocks on them here.
    // Completion is tracked by an mbarrier (cp.async.mbarrier.arrive.noinc): only the compute warp waits for
    // it; the streaming warps go straight to the history shift, which depends on none of the staged data.
    // Table-driven (one rolled loop) to keep the instruction footprint small: a CTA executes this code once,
    // cold, so straight-line code costs an instruction-cache miss per 128 bytes.
    ETRACE(0);
    if (tid == 0) {
        mbar_init(&S.mbar, T);
        S.cnt = 0;
        S.is_last = 0;
    }
    if (tid < HG_NUM_REWARDS) S.acc[tid] = 0.0f;
    // tile loop (grid-stride over 32-env tiles; with the default grid each CTA runs exactly one tile)
    unsigned tile_parity = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tile_parity ^= 1u) {
    const int e0 = tile * E;
    const int nE = min(E, N - e0);
    if (l2_prefetch && do_obs && tid == 0) {
        // The tile's history rows are one contiguous run of nE * pitch floats per buffer: ask L2 for all of it now (two bulk
        // prefetches, no registers, no shared memory), so that the register-staged shift below finds the rows on chip
        // instead of paying a DRAM round trip per batch of 21 requests.
        const uint32_t ob = (uint32_t)(nE * opitch * 4), pb = (uint32_t)(nE * ppitch * 4);
        const float* op = B.obs_buf + (size_t)e0 * opitch;
        const float* pp = B.privileged_obs_buf + (size_t)e0 * ppitch;
        if (((reinterpret_cast<uintptr_t>(op) | ob) & 15u) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(op), "r"(ob) : "memory");
        if (((reinterpret_cast<uintptr_t>(pp) | pb) & 15u) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pp), "r"(pb) : "memory");
    }
    if (tid < E) { S.reset[tid] = 0; S.root_dirty[tid] = 0; }
    if (tid == 0) { S.next_row[0] = 0; S.next_row[1] = 0; S.next_noise = 0; }
    __syncthreads();
    {
        const float* const* fields = reinterpret_cast<const float* const*>(&B);
        float* sbase = reinterpret_cast<float*>(&S);
        // one tile per WARP at a time: each tile costs a chain of dependent constant-memory look-ups (cold), which would be
        // paid 18 times in sequence if every thread walked every tile; spread over the warps the chains overlap
#pragma unroll 1
        for (int t = warp; t < kNumTiles; t += T / 32) {
            int w = kTileWidth[t];
            tile_load(sbase + kTileSmem[t], fields[kTileField[t]] + (size_t)e0 * w, nE * w, lane, 32);
        }
    }
#pragma unroll 1
    for (int i = tid; i < nE * 52; i += T) {              // feet / knee rows of rigid_state (52-byte runs)
        int le = i / 52, r = i - le * 52, b = r / 13, c = r - b * 13;
        int body = (b < 2) ? cP.feet[b] : cP.knees[b - 2];
        cp_async4(S.rg + i, B.rigid_state + ((size_t)(e0 + le) * nb + body) * 13 + c);
    }
#pragma unroll 1
    for (int i = tid; i < HG_NUM_REWARDS * E; i += T) {                // episode sums are (22, N): 128-byte runs
        int le = i % E;
        if (le < nE) cp_async4(S.sums + i, B.episode_sums + (size_t)(i / E) * N + e0 + le);
    }
    const int n_slots = 2 + cP.n_term + cP.n_pen;                      // contact-force rows the step reads
#pragma unroll 1
    for (int i = tid; i < nE * n_slots * 3; i += T) {
        int le = i / (n_slots * 3), r = i - le * (n_slots * 3), sl = r / 3, c = r - sl * 3;
        int body = sl < 2 ? cP.feet[sl] : (sl < 2 + cP.n_term ? cP.term_bodies[sl - 2] : cP.pen_bodies[sl - 2 - cP.n_term]);
        cp_async4(S.cf + le * (HG_CF_SLOTS * 3) + sl * 3 + c, B.contact_forces + ((size_t)(e0 + le) * nb + body) * 3 + c);
    }
    mbar_arrive_on_cp_async(&S.mbar);
    ETRACE(1);

    if (warp < NCW) {
        // ---- 2a. compute warps: one lane per env, FOUR warps per 32-env tile, each warp one "role" = a quarter of the per-env step.
        // The step is ~2.5 k useful instructions on a mostly dependent chain, executed cold (every SM fetches the code from L2 once
        // per launch, ~40 cycles per instruction): run by one warp it is the critical path of the kernel at every size (round 1:
        // ~13 us).  Roles are WARPS, not lanes -- a warp whose lanes took different roles would execute all four code paths one
        // after the other -- so each warp runs only its quarter of the code, on its own scheduler, and the quarters trade results
        // through the small `X` exchange block with 128-thread named barriers.  Whole sub-computations move, never parts of a
        // sum: every fp32 operation still happens in the reference's order.
        //   role 0: base-frame velocities / gravity (3 quaternion rotations), pitch, rewards 0 1 3 9, the ordered reward total
        //   role 1: roll, sin / cos of the gait clock, the 12-DoF reward loop and rewards 4 5 6 13 17
        //   role 2: yaw, the feet / knee rewards 2 7 8 10 11 12 14 and their state
        //   role 3: termination, command resampling / heading / push, rewards 15 16 18 19 20 21
        // Reset and observation assembly split per DoF (3 joints per role).
        const int role = warp;
        const int le = lane;
        const int e = e0 + le;
        const bool active = le < nE;
        const int lc_ = active ? le : 0;                         // inactive lanes read env 0's tiles and write nothing
        auto cbar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); };
        if (active && role == 0) {
            S.ep[le] = B.episode_length_buf[e];
            S.reset_in[le] = B.reset_buf[e];
            S.lc[2 * le] = B.last_contacts[(size_t)e * 2];
            S.lc[2 * le + 1] = B.last_contacts[(size_t)e * 2 + 1];
        }
        mbar_wait(&S.mbar, tile_parity);
        cbar();
        ETRACE(2);

        float* root = S.root + lc_ * 13;
        float* dof = S.dof + lc_ * 24;
        float* act = S.act + lc_ * 12;
        const float* cf = S.cf + lc_ * (HG_CF_SLOTS * 3);        // slots: feet L, feet R, termination bodies, penalised bodies
        const float* fL = S.rg + lc_ * 52;
        const float* fR = fL + 13;
        long long ep = S.ep[lc_];
        if (phases & HG_PHASE_COUNTERS) ep += 1;                  // legged_robot.py:128

        // ---- stage 1: base-frame quantities, gait clock, termination, callback ------------------------------------------
        if (phases & HG_PHASE_COUNTERS) {                         // legged_robot.py:129-136
            const float* q = root + 3;
            if (role == 0) {
                V3 v = quat_rotate_inverse(q, V3{root[7], root[8], root[9]});
                X.blv[0][le] = v.x; X.blv[1][le] = v.y; X.blv[2][le] = v.z;
                v = quat_rotate_inverse(q, V3{root[10], root[11], root[12]});
                X.bav[0][le] = v.x; X.bav[1][le] = v.y; X.bav[2][le] = v.z;
                v = quat_rotate_inverse(q, V3{0.0f, 0.0f, -1.0f});
                X.pg[0][le] = v.x; X.pg[1][le] = v.y; X.pg[2][le] = v.z;
            }
            if (role < 3) {                                       // get_euler_xyz + fold (:50-55): pitch / roll / yaw by roles 0 / 1 / 2
                float a;
                if (role == 0) {
                    float sinp = 2.0f * (q[3] * q[1] - q[2] * q[0]);
                    a = (fabsf(sinp) >= 1.0f) ? copysignf(1.5707963267948966f, sinp) : asinf(sinp);
                } else if (role == 1) {
                    a = atan2_call(2.0f * (q[3] * q[0] + q[1] * q[2]), q[3] * q[3] - q[0] * q[0] - q[1] * q[1] + q[2] * q[2]);
                } else {
                    a = atan2_call(2.0f * (q[3] * q[2] + q[0] * q[1]), q[3] * q[3] + q[0] * q[0] - q[1] * q[1] - q[2] * q[2]);
                }
                a = remainder_pos(a, kTwoPi);
                if (a > kPi) a -= kTwoPi;
                X.eul[role == 0 ? 1 : (role == 1 ? 0 : 2)][le] = a;
            }
        } else if (role == 0) {
            const int eg = active ? e : e0;
            const float* p = B.base_lin_vel + (size_t)eg * 3; X.blv[0][le] = p[0]; X.blv[1][le] = p[1]; X.blv[2][le] = p[2];
            p = B.base_ang_vel + (size_t)eg * 3; X.bav[0][le] = p[0]; X.bav[1][le] = p[1]; X.bav[2][le] = p[2];
            p = B.projected_gravity + (size_t)eg * 3; X.pg[0][le] = p[0]; X.pg[1][le] = p[1]; X.pg[2][le] = p[2];
            p = B.base_euler_xyz + (size_t)eg * 3;
            V3 eu{p[0], p[1], p[2]};
            if (do_reset) eu = euler_xyz_wrapped(root + 3);       // stand-alone reset_idx refreshes the euler angles of ALL envs (:213)
            X.eul[0][le] = eu.x; X.eul[1][le] = eu.y; X.eul[2][le] = eu.z;
        }
        // gait clock (humanoid_env.py:100-103): sin by role 1, cos by role 2, once per step -- the observation frames use the same
        // phase unless the env resets, and then it is exactly 0 (sin 0 = 0, cos 0 = 1)
        if (role == 1 && (phases & (HG_PHASE_REWARD | HG_PHASE_OBS))) X.s_r[le] = sinf(kTwoPi * ((float)ep * cP.dt / cP.cycle_time));
        if (role == 2 && do_obs) X.c_r[le] = cosf(kTwoPi * ((float)ep * cP.dt / cP.cycle_time));
        if (role == 3) {
            float cmd[4] = {S.in_a.cmd[lc_ * 4], S.in_a.cmd[lc_ * 4 + 1], S.in_a.cmd[lc_ * 4 + 2], S.in_a.cmd[lc_ * 4 + 3]};
            bool reset = (phases & HG_PHASE_COUNTERS) ? false : (S.reset_in[lc_] != 0);
            if (phases & HG_PHASE_TERMINATE) {                    // legged_robot.py:156-161
#pragma unroll 1
                for (int b = 0; b < cP.n_term; ++b) {
                    const float* f = cf + (2 + b) * 3;
                    reset |= sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 1.0f;
                }
                const bool timeout = ep > cP.max_episode_length;
                reset |= timeout;
                if (active) B.time_out_buf[e] = timeout;
            }
            X.reset_any[le] = reset ? 1 : 0;
            if (active) S.reset[le] = (do_reset && reset) ? 1 : 0;
            if (active && (phases & HG_PHASE_CALLBACK)) {         // legged_robot.py:304-320
                if (ep % cP.resample_period == 0) {
                    float u0 = draw_u(Z.u_cmd_cb, (int64_t)e * 3 + 0, Z.seed, Z.step, e, HG_RNG_CMD_CB, 0);
                    float u1 = draw_u(Z.u_cmd_cb, (int64_t)e * 3 + 1, Z.seed, Z.step, e, HG_RNG_CMD_CB, 1);
                    float u2 = draw_u(Z.u_cmd_cb, (int64_t)e * 3 + 2, Z.seed, Z.step, e, HG_RNG_CMD_CB, 2);
                    resample_commands(cP, cmd, u0, u1, u2);
                }
                if (cP.heading_command) {
                    // forward = quat_apply(q, (1,0,0)) = v + w t + u x t,  t = 2 (u x v)
                    float ux = root[3], uy = root[4], uz = root[5], w = root[6];
                    float tx = (uy * 0.0f - uz * 0.0f) * 2.0f, ty = (uz * 1.0f - ux * 0.0f) * 2.0f, tz = (ux * 0.0f - uy * 1.0f) * 2.0f;
                    float fx = 1.0f + w * tx + (uy * tz - uz * ty);
                    float fy = 0.0f + w * ty + (uz * tx - ux * tz);
                    float heading = atan2_call(fy, fx);
                    float a = remainder_pos(cmd[3] - heading, kTwoPi);          // utils/math.py:47-50
                    a = a - kTwoPi * (a > kPi ? 1.0f : 0.0f);
                    cmd[2] = clampf(0.5f * a, -1.0f, 1.0f);
                }
                if (cP.push_robots && (common_step % cP.push_interval == 0)) {   // humanoid_env.py:83-98
                    float u[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) u[k] = draw_u(Z.u_push, (int64_t)e * 5 + k, Z.seed, Z.step, e, HG_RNG_PUSH, k);
                    float f0 = cP.push_vel_span * u[0] + cP.push_vel_lo, f1 = cP.push_vel_span * u[1] + cP.push_vel_lo;
                    float t0 = cP.push_ang_span * u[2] + cP.push_ang_lo, t1 = cP.push_ang_span * u[3] + cP.push_ang_lo;
                    float t2 = cP.push_ang_span * u[4] + cP.push_ang_lo;
                    S.rpf[le * 3] = f0; S.rpf[le * 3 + 1] = f1;
                    S.rpt[le * 3] = t0; S.rpt[le * 3 + 1] = t1; S.rpt[le * 3 + 2] = t2;
                    float* rpf = B.rand_push_force + (size_t)e * 3;
                    float* rpt = B.rand_push_torque + (size_t)e * 3;
                    rpf[0] = f0; rpf[1] = f1;
                    rpt[0] = t0; rpt[1] = t1; rpt[2] = t2;
                    root[7] = f0; root[8] = f1;
                    root[10] = t0; root[11] = t1; root[12] = t2;
                    S.root_dirty[le] = 1;
                }
            }
            X.cmd[0][le] = cmd[0]; X.cmd[1][le] = cmd[1]; X.cmd[2][le] = cmd[2]; X.cmd[3][le] = cmd[3];
        }
        if (role == 2) { X.fat[0][le] = S.in_a.fat[2 * lc_]; X.fat[1][le] = S.in_a.fat[2 * lc_ + 1]; }
        cbar();                                                   // X.{blv,bav,pg,eul,s_r,cmd,reset_any,fat} and the push rewrite of root are visible

        const V3 blv{X.blv[0][le], X.blv[1][le], X.blv[2][le]}, bav{X.bav[0][le], X.bav[1][le], X.bav[2][le]};
        float cmd[4] = {X.cmd[0][le], X.cmd[1][le], X.cmd[2][le], X.cmd[3][le]};
        const bool reset = X.reset_any[le] != 0;
        bool cmd_dirty = (phases & HG_PHASE_CALLBACK) != 0;
        // feet contacts used by rewards and observations
        const bool contact0 = cf[2] > 5.0f, contact1 = cf[5] > 5.0f;

        // ---- stage 2: rewards (legged_robot.py:217-235): each role its terms, role 0 adds them up in alphabetical order ----
        if (phases & HG_PHASE_REWARD) {
            const float s = X.s_r[le];
            float st0 = s >= 0.0f ? 1.0f : 0.0f, st1 = s < 0.0f ? 1.0f : 0.0f;   // humanoid_env.py:105-118
            if (fabsf(s) < 0.1f) { st0 = 1.0f; st1 = 1.0f; }
            auto term = [&](int k, float rv) {                    // this role's scaled term: exchange slot, episode sum, optional per-term output
                const float v = rv * cP.reward_scales[k];
                X.rk[k][le] = v;
                if (active) {
                    S.sums[k * E + le] += v;
                    if (B.rew_terms) B.rew_terms[(size_t)k * N + e] = v;
                }
            };
            if (role == 0) {
                const float* lact = S.in_a.lact + lc_ * 12;
                const float* llact = S.in_a.llact + lc_ * 12;
                {   // 0 action_smoothness, humanoid_env.py:530-540
                    float t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
#pragma unroll 1
                    for (int j = 0; j < 12; ++j) {
                        float d1 = lact[j] - act[j];
                        t1 += d1 * d1;
                        float d2 = act[j] + llact[j] - 2.0f * lact[j];
                        t2 += d2 * d2;
                        t3 += fabsf(act[j]);
                    }
                    term(0, t1 + t2 + 0.05f * t3);
                }
                {   // 1 base_acc :386-393
                    const float* lrv = S.in_a.lrv + lc_ * 6;
                    float a = 0.0f;
                    for (int j = 0; j < 6; ++j) { float d = lrv[j] - root[7 + j]; a += d * d; }
                    term(1, expf(-sqrtf(a) * 3.0f));
                }
                {   // 3 collision :523-528
                    float c = 0.0f;
#pragma unroll 1
                    for (int b = 0; b < cP.n_pen; ++b) {
                        const float* f = cf + (2 + cP.n_term + b) * 3;
                        c += (sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 0.1f) ? 1.0f : 0.0f;
                    }
                    term(3, c);
                }
                {   // 9 feet_contact_forces :355-360
                    const float* a = cf;
                    const float* b = cf + 3;
                    float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
                    float nbn = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
                    term(9, clampf(na - cP.max_contact_force, 0.0f, 400.0f) + clampf(nbn - cP.max_contact_force, 0.0f, 400.0f));
                }
            } else if (role == 1) {
                const float* ldv = S.in_a.ldv + lc_ * 12;
                const float* tau = S.in_a.tau + lc_ * 12;
                const float* ref = S.in_a.ref + lc_ * 12;        // STALE reference pose (hazard 2)
                float dq2 = 0.0f, dacc = 0.0f, qerr = 0.0f, jall = 0.0f, tq = 0.0f;
#pragma unroll 1
                for (int j = 0; j < 12; ++j) {
                    float q = dof[2 * j], v = dof[2 * j + 1];
                    float d = q - cP.default_dof_pos[j];
                    jall += d * d;
                    dq2 += v * v;
                    float a = (ldv[j] - v) / cP.dt;
                    dacc += a * a;
                    float er = q - ref[j];
                    qerr += er * er;
                    tq += tau[j] * tau[j];
                }
                {   // 4 default_joint_pos :362-372
                    float d0 = dof[0] - cP.default_dof_pos[0], d1 = dof[2] - cP.default_dof_pos[1];
                    float d6 = dof[12] - cP.default_dof_pos[6], d7 = dof[14] - cP.default_dof_pos[7];
                    float yr = sqrtf(d0 * d0 + d1 * d1) + sqrtf(d6 * d6 + d7 * d7);
                    yr = clampf(yr - 0.1f, 0.0f, 50.0f);
                    term(4, expf(-yr * 100.0f) - 0.01f * sqrtf(jall));
                }
                term(5, dacc);                                    // 5 dof_acc :516-521
                term(6, dq2);                                     // 6 dof_vel :509-514
                {   // 13 joint_pos :272-280
                    float er = sqrtf(qerr);
                    term(13, expf(-2.0f * er) - 0.2f * clampf(er, 0.0f, 0.5f));
                }
                term(17, tq);                                     // 17 torques :502-507
            } else if (role == 2) {
                float fat0 = X.fat[0][le], fat1 = X.fat[1][le];
                {   // 2 base_height :374-384
                    float measured = (fL[2] * st0 + fR[2] * st1) / (st0 + st1);
                    float h = root[2] - (measured - 0.05f);
                    term(2, expf(-fabsf(h - cP.base_height_target) * 100.0f));
                }
                {   // 7 feet_air_time :320-334 (stateful)
                    bool filt0 = contact0 || (st0 != 0.0f) || S.lc[2 * lc_];
                    bool filt1 = contact1 || (st1 != 0.0f) || S.lc[2 * lc_ + 1];
                    if (active) {
                        B.last_contacts[(size_t)e * 2] = contact0;
                        B.last_contacts[(size_t)e * 2 + 1] = contact1;
                    }
                    float a0 = fat0, a1 = fat1;
                    bool first0 = (a0 > 0.0f) && filt0, first1 = (a1 > 0.0f) && filt1;
                    a0 += cP.dt; a1 += cP.dt;
                    term(7, clampf(a0, 0.0f, 0.5f) * (first0 ? 1.0f : 0.0f) + clampf(a1, 0.0f, 0.5f) * (first1 ? 1.0f : 0.0f));
                    fat0 = a0 * (filt0 ? 0.0f : 1.0f);
                    fat1 = a1 * (filt1 ? 0.0f : 1.0f);
                    X.fat[0][le] = fat0; X.fat[1][le] = fat1;
                }
                {   // 8 feet_clearance :446-467 (stateful; never reset, hazard 4)
                    float z0 = fL[2] - 0.05f, z1 = fR[2] - 0.05f;
                    float h0 = S.in_a.fh[2 * lc_] + (z0 - S.in_a.lfz[2 * lc_]), h1 = S.in_a.fh[2 * lc_ + 1] + (z1 - S.in_a.lfz[2 * lc_ + 1]);
                    float hit0 = fabsf(h0 - cP.target_feet_height) < 0.01f ? 1.0f : 0.0f;
                    float hit1 = fabsf(h1 - cP.target_feet_height) < 0.01f ? 1.0f : 0.0f;
                    term(8, hit0 * (1.0f - st0) + hit1 * (1.0f - st1));
                    if (active) {
                        float2* gh = reinterpret_cast<float2*>(B.feet_height) + e;
                        float2* gz = reinterpret_cast<float2*>(B.last_feet_z) + e;
                        *gh = make_float2(h0 * (contact0 ? 0.0f : 1.0f), h1 * (contact1 ? 0.0f : 1.0f));
                        *gz = make_float2(z0, z1);
                    }
                }
                {   // 10 feet_contact_number :336-344
                    float m0 = ((contact0 ? 1.0f : 0.0f) == st0) ? 1.0f : -0.3f;
                    float m1 = ((contact1 ? 1.0f : 0.0f) == st1) ? 1.0f : -0.3f;
                    term(10, (m0 + m1) / 2.0f);
                }
                term(11, two_point_distance_reward(fL[0], fL[1], fR[0], fR[1], cP.min_dist, cP.max_dist));   // 11 :282-292
                {   // 12 foot_slip :308-318
                    float v0 = sqrtf(sqrtf(fL[7] * fL[7] + fL[8] * fL[8]));
                    float v1 = sqrtf(sqrtf(fR[7] * fR[7] + fR[8] * fR[8]));
                    term(12, v0 * (contact0 ? 1.0f : 0.0f) + v1 * (contact1 ? 1.0f : 0.0f));
                }
                {   // 14 knee_distance :295-305
                    const float* kL = fL + 26;
                    const float* kR = fL + 39;
                    term(14, two_point_distance_reward(kL[0], kL[1], kR[0], kR[1], cP.min_dist, cP.max_dist / 2.0f));
                }
            } else {
                const V3 pg{X.pg[0][le], X.pg[1][le], X.pg[2][le]}, eul{X.eul[0][le], X.eul[1][le], X.eul[2][le]};
                {   // 15 low_speed :469-500
                    float v = blv.x, c = cmd[0];
                    float av = fabsf(v), ac = fabsf(c);
                    bool low = av < 0.5f * ac, high = av > 1.2f * ac;
                    float rr = 0.0f;
                    if (low) rr = -1.0f;
                    if (high) rr = 0.0f;
                    if (!(low || high)) rr = 1.2f;
                    float sv = (v > 0.0f) - (v < 0.0f), sc = (c > 0.0f) - (c < 0.0f);
                    if (sv != sc) rr = -2.0f;
                    term(15, rr * (ac > 0.1f ? 1.0f : 0.0f));
                }
                {   // 16 orientation :346-353
                    float a = expf(-(fabsf(eul.x) + fabsf(eul.y)) * 10.0f);
                    float b = expf(-sqrtf(pg.x * pg.x + pg.y * pg.y) * 20.0f);
                    term(16, (a + b) / 2.0f);
                }
                {
                    float ex = cmd[0] - blv.x, ey = cmd[1] - blv.y;
                    float le2 = ex * ex + ey * ey;
                    float lin = sqrtf(le2);
                    float ang = fabsf(cmd[2] - bav.z);
                    term(18, (expf(-lin * 10.0f) + expf(-ang * 10.0f)) / 2.0f - 0.2f * (lin + ang));   // 18 track_vel_hard :408-425
                    float da = cmd[2] - bav.z;
                    term(19, expf(-(da * da) * cP.tracking_sigma));                                    // 19 tracking_ang_vel :436-444
                    term(20, expf(-le2 * cP.tracking_sigma));                                          // 20 tracking_lin_vel :427-434
                }
                {   // 21 vel_mismatch_exp :396-406
                    float a = expf(-(blv.z * blv.z) * 10.0f);
                    float b = expf(-sqrtf(bav.x * bav.x + bav.y * bav.y) * 5.0f);
                    term(21, (a + b) / 2.0f);
                }
            }
        }
        cbar();                        // every role's terms are in X.rk; stage 2's reads of act / dof / root / sums are over (stage 3 rewrites them)
        ETRACE(3);
        if ((phases & HG_PHASE_REWARD) && role == 0) {            // alphabetical accumulation (legged_robot.py:222-230)
            float total = 0.0f;
#pragma unroll
            for (int k = 0; k < HG_NUM_REWARDS; ++k) total += X.rk[k][le];
            if (cP.only_positive_rewards) total = fmaxf(total, 0.0f);
            if (active) B.rew_buf[e] = total;
        }

        // ---- stage 3: reset_idx for the envs that terminated (legged_robot.py:163-215) ----------------------------------
        const bool rz = do_reset && reset;
        if (rz && active) {
#pragma unroll 1
            for (int j = 3 * role; j < 3 * role + 3; ++j) {     // _reset_dofs :359-373, three joints per role
                float u = draw_u(Z.u_dof, (int64_t)e * 12 + j, Z.seed, Z.step, e, HG_RNG_DOF, j);
                dof[2 * j] = cP.default_dof_pos[j] + (cP.dof_reset_span * u + cP.dof_reset_lo);
                dof[2 * j + 1] = 0.0f;
                act[j] = 0.0f;
            }
            if (role == 0) {
                for (int j = 0; j < 13; ++j) root[j] = cP.base_init_state[j];      // _reset_root_states :374-397
                root[0] += S.org[le * 3]; root[1] += S.org[le * 3 + 1]; root[2] += S.org[le * 3 + 2];
                S.root_dirty[le] = 1;
                atomicAdd(&S.cnt, 1);
                B.reset_ids[atomicAdd(&B.scratch[0], 1)] = e;
                const V3 eu = euler_xyz_wrapped(root + 3);        // "fix reset gravity bug" :212-215
                X.eul[0][le] = eu.x; X.eul[1][le] = eu.y; X.eul[2][le] = eu.z;
                const V3 g3 = quat_rotate_inverse(root + 3, V3{0.0f, 0.0f, -1.0f});
                X.pg[0][le] = g3.x; X.pg[1][le] = g3.y; X.pg[2][le] = g3.z;
            }
            if (role == 3) {
                float u0 = draw_u(Z.u_cmd_rs, (int64_t)e * 3 + 0, Z.seed, Z.step, e, HG_RNG_CMD_RS, 0);
                float u1 = draw_u(Z.u_cmd_rs, (int64_t)e * 3 + 1, Z.seed, Z.step, e, HG_RNG_CMD_RS, 1);
                float u2 = draw_u(Z.u_cmd_rs, (int64_t)e * 3 + 2, Z.seed, Z.step, e, HG_RNG_CMD_RS, 2);
                resample_commands(cP, cmd, u0, u1, u2);
                X.cmd[0][le] = cmd[0]; X.cmd[1][le] = cmd[1]; X.cmd[2][le] = cmd[2]; X.cmd[3][le] = cmd[3];
            }
#pragma unroll 1
            for (int k = role; k < HG_NUM_REWARDS; k += NCW) {  // extras["episode"] :198-202
                atomicAdd(&S.acc[k], S.sums[k * E + le]);
                S.sums[k * E + le] = 0.0f;
            }
            if (role == 2) { X.fat[0][le] = 0.0f; X.fat[1][le] = 0.0f; }
            ep = 0;
            cmd_dirty = true;
        }
        if (active && role == 0 && (phases & (HG_PHASE_TERMINATE | HG_PHASE_RESET))) B.reset_buf[e] = reset;
        if (active && role == 1 && do_last) {   // last_last_actions <- last_actions (0 if reset), legged_robot.py:147 (+ :190); its tile is about to be reused
            const float4* la = reinterpret_cast<const float4*>(S.in_a.lact + le * 12);
            float4* dst = reinterpret_cast<float4*>(B.last_last_actions + (size_t)e * 12);
            const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            dst[0] = rz ? z4 : la[0]; dst[1] = rz ? z4 : la[1]; dst[2] = rz ? z4 : la[2];
        }
        // every compute warp is done with the aliased input tiles (in_a, rg) before newpriv / newobs are written over them;
        // the reset rewrites (dof, act, root, X.cmd / eul / pg / fat) and X.sc_obs are visible
        cbar();
        ETRACE(4);
        const V3 pg{X.pg[0][le], X.pg[1][le], X.pg[2][le]}, eul{X.eul[0][le], X.eul[1][le], X.eul[2][le]};
        cmd[0] = X.cmd[0][le]; cmd[1] = X.cmd[1][le]; cmd[2] = X.cmd[2][le]; cmd[3] = X.cmd[3][le];

        // ---- stage 4: observation frames (humanoid_env.py:200-262), three joints and a share of the scalars per role -----
        if (do_obs && active) {
            const float s = rz ? 0.0f : X.s_r[le], c = rz ? 1.0f : X.c_r[le];
            float st0 = s >= 0.0f ? 1.0f : 0.0f, st1 = s < 0.0f ? 1.0f : 0.0f;
            bool dbl = fabsf(s) < 0.1f;
            if (dbl) { st0 = 1.0f; st1 = 1.0f; }
            float sl = s > 0.0f ? 0.0f : s, sr = s < 0.0f ? 0.0f : s;   // compute_ref_state :121-142
            float k1 = cP.target_joint_pos_scale, k2 = 2.0f * k1;
            float* o = S.newobs + le * HG_OBS1;
            float* p = S.newpriv + le * HG_PRIV1;
            float* gref = B.ref_dof_pos + (size_t)e * 12;
#pragma unroll 1
            for (int j = 3 * role; j < 3 * role + 3; ++j) {
                float rj = 0.0f;
                if (!dbl) {
                    if (j == 2 || j == 4) rj = sl * k1;
                    if (j == 3) rj = sl * k2;
                    if (j == 8 || j == 10) rj = sr * k1;
                    if (j == 9) rj = sr * k2;
                }
                float q = dof[2 * j], v = dof[2 * j + 1];
                float qs = (q - cP.default_dof_pos[j]) * cP.obs_scale_dof_pos;
                float vs = v * cP.obs_scale_dof_vel;
                o[5 + j] = qs; o[17 + j] = vs; o[29 + j] = act[j];
                p[5 + j] = qs; p[17 + j] = vs; p[29 + j] = act[j];
                p[41 + j] = q - rj;
                gref[j] = rj;
            }
            if (role == 0) {
                float ci[5] = {s, c, cmd[0] * cP.obs_scale_lin_vel, cmd[1] * cP.obs_scale_lin_vel, cmd[2] * cP.obs_scale_ang_vel};
#pragma unroll
                for (int j = 0; j < 5; ++j) { o[j] = ci[j]; p[j] = ci[j]; }
            } else if (role == 1) {
                float a0 = bav.x * cP.obs_scale_ang_vel, a1 = bav.y * cP.obs_scale_ang_vel, a2 = bav.z * cP.obs_scale_ang_vel;
                float e0_ = eul.x * cP.obs_scale_quat, e1 = eul.y * cP.obs_scale_quat, e2 = eul.z * cP.obs_scale_quat;
                o[41] = a0; o[42] = a1; o[43] = a2; o[44] = e0_; o[45] = e1; o[46] = e2;
                p[56] = a0; p[57] = a1; p[58] = a2; p[59] = e0_; p[60] = e1; p[61] = e2;
            } else if (role == 2) {
                p[53] = blv.x * cP.obs_scale_lin_vel; p[54] = blv.y * cP.obs_scale_lin_vel; p[55] = blv.z * cP.obs_scale_lin_vel;
                p[62] = S.rpf[le * 3]; p[63] = S.rpf[le * 3 + 1];
                p[64] = S.rpt[le * 3]; p[65] = S.rpt[le * 3 + 1]; p[66] = S.rpt[le * 3 + 2];
                p[67] = S.fric[le];
                p[68] = S.mass[le] / 30.0f;
            } else {
                p[69] = st0; p[70] = st1;
                p[71] = contact0 ? 1.0f : 0.0f; p[72] = contact1 ? 1.0f : 0.0f;
            }
        }

        // ---- stage 5: small per-env outputs (fire-and-forget stores), a few per role ---------------------------------
        if (active) {
            if (role == 0 && (phases & (HG_PHASE_COUNTERS | HG_PHASE_RESET))) {
                B.episode_length_buf[e] = ep;
                float* q;
                q = B.projected_gravity + (size_t)e * 3; q[0] = pg.x; q[1] = pg.y; q[2] = pg.z;
                q = B.base_euler_xyz + (size_t)e * 3; q[0] = eul.x; q[1] = eul.y; q[2] = eul.z;
            }
            if (role == 1 && (phases & HG_PHASE_COUNTERS)) {
                float* q;
                q = B.base_lin_vel + (size_t)e * 3; q[0] = blv.x; q[1] = blv.y; q[2] = blv.z;
                q = B.base_ang_vel + (size_t)e * 3; q[0] = bav.x; q[1] = bav.y; q[2] = bav.z;
            }
            if (role == 2 && (phases & (HG_PHASE_REWARD | HG_PHASE_RESET)))
                *(reinterpret_cast<float2*>(B.feet_air_time) + e) = make_float2(X.fat[0][le], X.fat[1][le]);
            if (role == 3 && cmd_dirty) *reinterpret_cast<float4*>(B.commands + (size_t)e * 4) = make_float4(cmd[0], cmd[1], cmd[2], cmd[3]);
            if (role == 3 && do_last) {                           // legged_robot.py:150
                float* lrv = B.last_root_vel + (size_t)e * 6;
#pragma unroll
                for (int j = 0; j < 6; ++j) lrv[j] = root[7 + j];
            }
        }
    }
    ETRACE(5);
    // ---- 2b. history shift: every warp pulls rows off a per-tile queue (the compute warps join when they are done), overlapped
    // with staging and the reward / observation math.  Rows are copied unconditionally; those of envs that reset are zeroed in step 3.
    if (do_obs) {
        stream_history<HG_OBS1, OBS_KEEP, HIST_OBS_ROWS>(B.obs_out + (size_t)e0 * opitch, B.obs_buf + (size_t)e0 * opitch, &S.next_row[0], nE, lane, opitch);
        stream_history<HG_PRIV1, PRIV_KEEP, 4>(B.priv_out + (size_t)e0 * ppitch, B.privileged_obs_buf + (size_t)e0 * ppitch, &S.next_row[1],
                                               nE, lane, ppitch);
    }
    // ---- 2c. observation noise (humanoid_env.py:249-252): the draws depend on nothing computed above, so whichever warps get here
    // first (the streaming warps of a wide CTA) produce them while the compute warps are still busy.  One Philox4x32-10 call
    // yields the Box-Muller pair of TWO channels; chunks of 32 channel pairs come off a queue.
    if (do_obs && cP.add_noise) {
        constexpr int PAIRS = (HG_OBS1 + 1) / 2;                  // 24
#pragma unroll 1
        for (;;) {
            int c0 = 0;
            if (lane == 0) c0 = atomicAdd(&S.next_noise, 32);
            c0 = __shfl_sync(0xffffffffu, c0, 0);
            if (c0 >= nE * PAIRS) break;
            const int i = c0 + lane;
            if (i < nE * PAIRS) {
                const int le = i / PAIRS, k = 2 * (i - le * PAIRS);
                const int e = e0 + le;
                float z0, z1 = 0.0f;
                if (Z.z_obs) {
                    z0 = Z.z_obs[(size_t)e * HG_OBS1 + k];
                    if (k + 1 < HG_OBS1) z1 = Z.z_obs[(size_t)e * HG_OBS1 + k + 1];
                } else {
                    HgPhilox r = philox_call(Z.seed, (uint32_t)e, (uint32_t)Z.step, HG_RNG_OBS | ((uint32_t)(Z.step >> 32) << 8), (uint32_t)k);
                    const float u1 = ((float)(r.c[0] >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = hg_u01(r.c[1]);
                    const float rad = sqrtf(-2.0f * logf(u1));
                    float sn, cs;
                    sincospif(2.0f * u2, &sn, &cs);
                    z0 = rad * cs;
                    z1 = rad * sn;
                }
                S.z[le * HG_OBS1 + k] = z0;
                if (k + 1 < HG_OBS1) S.z[le * HG_OBS1 + k + 1] = z1;
            }
        }
    }
    ETRACE(6);
    __syncthreads();
    ETRACE(7);

    // ---- 3. newest frame: noise (humanoid_env.py:249-252), +-18 clip (legged_robot.py:104-108), store ---------
    if (do_obs) {
        for (int i = tid; i < nE * HG_OBS1; i += T) {
            int le = i / HG_OBS1, k = i - le * HG_OBS1;
            float v = S.newobs[i];
            if (cP.add_noise) {
                const float sc = cP.noise_scale_vec[k];
                if (sc != 0.0f || Z.z_obs) v = v + S.z[i] * sc * cP.noise_level;
            }
            if (do_last) v = clampf(v, -cP.clip_obs, cP.clip_obs);
            B.obs_out[(size_t)(e0 + le) * opitch + OBS_KEEP + k] = v;
        }
        for (int i = tid; i < nE * HG_PRIV1; i += T) {
            int le = i / HG_PRIV1, k = i - le * HG_PRIV1;
            float v = S.newpriv[i];
            if (do_last) v = clampf(v, -cP.clip_obs, cP.clip_obs);
            B.priv_out[(size_t)(e0 + le) * ppitch + PRIV_KEEP + k] = v;
        }
        if (do_reset) {   // reset envs restart with an all-zero history (humanoid_env.py:264-269); rare
#pragma unroll 1
            for (int le = 0; le < nE; ++le) {
                if (!S.reset[le]) continue;
                for (int k = tid; k < OBS_KEEP; k += T) B.obs_out[(size_t)(e0 + le) * opitch + k] = 0.0f;
                for (int k = tid; k < PRIV_KEEP; k += T) B.priv_out[(size_t)(e0 + le) * ppitch + k] = 0.0f;
            }
        }
    } else if (do_reset) {
        // stand-alone reset_idx: zero the history rows of the reset envs in place (humanoid_env.py:264-269)
        for (int i = tid; i < nE * OBS_W; i += T)
            if (S.reset[i / OBS_W]) B.obs_buf[(size_t)(e0 + i / OBS_W) * opitch + i % OBS_W] = 0.0f;
        for (int i = tid; i < nE * PRIV_W; i += T)
            if (S.reset[i / PRIV_W]) B.privileged_obs_buf[(size_t)(e0 + i / PRIV_W) * ppitch + i % PRIV_W] = 0.0f;
    }

    ETRACE(8);
    // ---- 4. coalesced write-back of the tiles the step modified ------------------------------------------------
    if (phases & (HG_PHASE_REWARD | HG_PHASE_RESET)) {
#pragma unroll 1
        for (int i = tid; i < HG_NUM_REWARDS * E; i += T) {
            int le = i % E;
            if (le < nE) B.episode_sums[(size_t)(i / E) * N + e0 + le] = S.sums[i];
        }
    }
    if (do_reset || (phases & HG_PHASE_CALLBACK)) {
        // root / dof rows change only on push or reset: write back the dirty rows
        for (int i = tid; i < nE * 13; i += T)
            if (S.root_dirty[i / 13]) B.root_states[(size_t)e0 * 13 + i] = S.root[i];
        for (int i = tid; i < nE * 24; i += T)
            if (S.reset[i / 24]) B.dof_state[(size_t)e0 * 24 + i] = S.dof[i];
    }
    if (do_last) {                                              // legged_robot.py:147-149 (+ reset zeroing :190-195)
        // last_last <- last (0 if reset); last <- actions (0 if reset); last_dof_vel <- dof_vel
        for (int i = tid; i < nE * 12; i += T) {
            int le = i / 12, j = i - le * 12;
            bool rz = S.reset[le];
            size_t gi = (size_t)e0 * 12 + i;
            B.last_actions[gi] = S.act[i];
            B.last_dof_vel[gi] = S.dof[le * 24 + 2 * j + 1];
            if (rz) B.actions[gi] = 0.0f;
        }
    } else if (do_reset) {
        for (int i = tid; i < nE * 12; i += T) {
            if (S.reset[i / 12]) {
                size_t gi = (size_t)e0 * 12 + i;
                B.last_last_actions[gi] = 0.0f; B.last_actions[gi] = 0.0f; B.actions[gi] = 0.0f; B.last_dof_vel[gi] = 0.0f;
            }
        }
    }

    __syncthreads();      // smem tiles are reused by the next tile of this CTA
    ETRACE(9);
    }                     // tile loop

    // ---- 5. extras["episode"] means + stale-able time_outs: last CTA finalises -------------------
    {
        if (S.cnt > 0 && tid < HG_NUM_REWARDS) atomicAdd(reinterpret_cast<float*>(B.scratch + 8) + tid, S.acc[tid]);
        __threadfence();
        __syncthreads();
        if (tid == 0) S.is_last = (atomicAdd(&B.scratch[1], 1) == (int)gridDim.x - 1);
        __syncthreads();
        if (S.is_last) {
            __threadfence();
            int total = *reinterpret_cast<volatile int*>(&B.scratch[0]);
            if (total > 0) {
                if (tid < HG_NUM_REWARDS) {
                    float sum = *(reinterpret_cast<volatile float*>(B.scratch + 8) + tid);
                    B.episode_means[tid] = sum / (float)total / cP.max_episode_length_s;
                }
                // N-byte copy by ONE CTA: 16-byte vectors, 4 independent requests per thread in flight
                if (((reinterpret_cast<uintptr_t>(B.extras_time_outs) | reinterpret_cast<uintptr_t>(B.time_out_buf)) & 15u) == 0) {
                    const uint4* src = reinterpret_cast<const uint4*>(B.time_out_buf);
                    uint4* dst = reinterpret_cast<uint4*>(B.extras_time_outs);
                    const int n16 = N >> 4;
#pragma unroll 4
                    for (int i = tid; i < n16; i += T) dst[i] = src[i];
                    for (int i = (n16 << 4) + tid; i < N; i += T) B.extras_time_outs[i] = B.time_out_buf[i];
                } else {
#pragma unroll 4
                    for (int i = tid; i < N; i += T) B.extras_time_outs[i] = B.time_out_buf[i];
                }
            }
            __syncthreads();
            if (tid < HG_NUM_REWARDS) reinterpret_cast<float*>(B.scratch + 8)[tid] = 0.0f;
            if (tid == 0) {
                if (do_reset) B.scratch[3] = total;
                B.scratch[0] = 0; B.scratch[1] = 0;
                if (Z.use_device_counters) {
                    if (phases & HG_PHASE_COUNTERS) *reinterpret_cast<int64_t*>(B.scratch + 4) += 1;
                    *reinterpret_cast<uint64_t*>(B.scratch + 6) += 1;
                }
            }
        }
    }
    ETRACE(10);
}

// ---------------------------------------------------------------------------------------------------
__global__ void pre_physics_kernel(const __grid_constant__ HgEnvParams cP, HgEnvBuffers B, const float* __restrict__ actions_in, const float* __restrict__ u_delay,
                                   const float* __restrict__ z_act, uint64_t seed, uint64_t step, int N) {
    if (step == ~0ull) step = *reinterpret_cast<const uint64_t*>(B.scratch + 6);   // device-side counter
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 12) return;
    int e = i / 12, j = i - e * 12;
    float c = cP.clip_actions;
    float a = clampf(actions_in[i], -c, c);                     // humanoid_env.py:192
    float u, z;
    if (u_delay) u = u_delay[e];
    else u = hg_u01(hg_philox(seed, e, (uint32_t)step, HG_RNG_DELAY | ((uint32_t)(step >> 32) << 8), 0).c[0]);
    if (z_act) z = z_act[i];
    else {
        HgPhilox r = hg_philox(seed, e, (uint32_t)step, HG_RNG_ACT | ((uint32_t)(step >> 32) << 8), j);
        z = hg_normal(r.c[0], r.c[1]);
    }
    float delay = u * cP.action_delay;                          // :194-195
    a = (1.0f - delay) * a + delay * B.actions[i];
    a = a + cP.action_noise * z * a;                            // :196
    B.actions[i] = clampf(a, -c, c);                            // legged_robot.py:90-91
}

__global__ void compute_torques_kernel(const __grid_constant__ HgEnvParams cP, HgEnvBuffers B, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 12) return;
    int j = i % 12;
    float2 s = reinterpret_cast<const float2*>(B.dof_state)[i];
    float scaled = B.actions[i] * cP.action_scale;
    float t = cP.p_gains[j] * (scaled + cP.default_dof_pos[j] - s.x) - cP.d_gains[j] * s.y;
    B.torques[i] = clampf(t, -cP.torque_limits[j], cP.torque_limits[j]);
}

// Synthetic-physics fast path of the decimation loop (legged_robot.py:94-101): with an open-loop frame source the
// `decimation` x {PD torque, set forces, simulate, refresh dof} sub-steps and the three state refreshes of
// post_physics_step (:124-126) are ONE launch instead of ~23 graph nodes.  Every sub-step's PD law is still
// evaluated against the dof state the previous sub-step left (the first one against the live state, which carries the
// reset rewrites); only the last torque and the last dof frame are observable downstream, exactly as in the loop.
__global__ void synth_decimation_kernel(const __grid_constant__ HgEnvParams cP, HgEnvBuffers B, const float2* __restrict__ dof_frames, int decimation,
                                        const float* __restrict__ root_f, const float* __restrict__ contact_f,
                                        const float* __restrict__ rigid_f, int N, int nb) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    const int nd = N * 12;
    float2* live = reinterpret_cast<float2*>(B.dof_state);
    for (int i = tid; i < nd; i += nt) {
        const int j = i % 12;
        float2 s = live[i];
        const float scaled = B.actions[i] * cP.action_scale;
        float t = 0.0f;
        for (int d = 0; d < decimation; ++d) {
            t = cP.p_gains[j] * (scaled + cP.default_dof_pos[j] - s.x) - cP.d_gains[j] * s.y;
            t = clampf(t, -cP.torque_limits[j], cP.torque_limits[j]);
            s = __ldg(dof_frames + (size_t)d * nd + i);
        }
        B.torques[i] = t;
        live[i] = s;
    }
    float* contact = const_cast<float*>(B.contact_forces);
    float* rigid = const_cast<float*>(B.rigid_state);
    for (int i = tid; i < N * 13; i += nt) B.root_states[i] = __ldg(root_f + i);
    for (int i = tid; i < N * nb * 3; i += nt) contact[i] = __ldg(contact_f + i);
    const int nr4 = (N * nb * 13) >> 2;
    if ((((uintptr_t)rigid | (uintptr_t)rigid_f) & 15u) == 0) {
        for (int i = tid; i < nr4; i += nt) reinterpret_cast<float4*>(rigid)[i] = __ldg(reinterpret_cast<const float4*>(rigid_f) + i);
        for (int i = (nr4 << 2) + tid; i < N * nb * 13; i += nt) rigid[i] = __ldg(rigid_f + i);
    } else {
        for (int i = tid; i < N * nb * 13; i += nt) rigid[i] = __ldg(rigid_f + i);
    }
}

long long* g_env_trace = nullptr;

int32_t check_buffers(const HgEnvBuffers* B) {
    HG_REQUIRE(B);
    HG_REQUIRE(B->root_states); HG_REQUIRE(B->dof_state); HG_REQUIRE(B->contact_forces); HG_REQUIRE(B->rigid_state);
    HG_REQUIRE(B->actions); HG_REQUIRE(B->last_actions); HG_REQUIRE(B->last_last_actions); HG_REQUIRE(B->torques);
    HG_REQUIRE(B->last_dof_vel); HG_REQUIRE(B->last_root_vel); HG_REQUIRE(B->commands); HG_REQUIRE(B->episode_length_buf);
    HG_REQUIRE(B->reset_buf); HG_REQUIRE(B->time_out_buf); HG_REQUIRE(B->extras_time_outs); HG_REQUIRE(B->base_lin_vel);
    HG_REQUIRE(B->base_ang_vel); HG_REQUIRE(B->projected_gravity); HG_REQUIRE(B->base_euler_xyz); HG_REQUIRE(B->feet_air_time);
    HG_REQUIRE(B->last_contacts); HG_REQUIRE(B->feet_height); HG_REQUIRE(B->last_feet_z); HG_REQUIRE(B->ref_dof_pos);
    HG_REQUIRE(B->rand_push_force); HG_REQUIRE(B->rand_push_torque); HG_REQUIRE(B->env_frictions); HG_REQUIRE(B->body_mass);
    HG_REQUIRE(B->env_origins); HG_REQUIRE(B->episode_sums); HG_REQUIRE(B->episode_means); HG_REQUIRE(B->obs_buf);
    HG_REQUIRE(B->privileged_obs_buf); HG_REQUIRE(B->rew_buf); HG_REQUIRE(B->reset_ids); HG_REQUIRE(B->scratch);
    HG_REQUIRE(B->obs_out); HG_REQUIRE(B->priv_out);
    if (!hg_aligned16(B->commands)) return hg_fail(HG_E_ALIGN, "commands must be 16-byte aligned");
    if (!hg_aligned16(B->dof_state)) return hg_fail(HG_E_ALIGN, "dof_state must be 16-byte aligned");
    return 0;
}

}  // namespace

extern "C" int32_t hg_env_pre_physics(const HgEnvBuffers* B, const HgEnvParams* P, const float* actions_in,
                                      const float* u_delay, const float* z_act, uint64_t seed, uint64_t step,
                                      int64_t N, void* stream) {
    HG_REQUIRE(B); HG_REQUIRE(P); HG_REQUIRE(actions_in); HG_REQUIRE(B->actions);
    if (N <= 0 || N > (1 << 26)) return hg_fail(HG_E_SIZE, "hg_env_pre_physics: bad N");
    cudaStream_t st = (cudaStream_t)stream;
    int total = (int)N * 12;
    pre_physics_kernel<<<(total + 255) / 256, 256, 0, st>>>(*P, *B, actions_in, u_delay, z_act, seed, step, (int)N);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_env_pre_physics");
}

extern "C" int32_t hg_env_compute_torques(const HgEnvBuffers* B, const HgEnvParams* P, int64_t N, void* stream) {
    HG_REQUIRE(B); HG_REQUIRE(P); HG_REQUIRE(B->actions); HG_REQUIRE(B->dof_state); HG_REQUIRE(B->torques);
    if (N <= 0 || N > (1 << 26)) return hg_fail(HG_E_SIZE, "hg_env_compute_torques: bad N");
    cudaStream_t st = (cudaStream_t)stream;
    int total = (int)N * 12;
    compute_torques_kernel<<<(total + 255) / 256, 256, 0, st>>>(*P, *B, (int)N);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_env_compute_torques");
}

// debug aid: the following hg_env_post_physics launches stamp %globaltimer at their phase boundaries into buf ([grid][12] int64)
extern "C" void hg_env_set_trace(long long* buf) { g_env_trace = buf; }

extern "C" int32_t hg_env_synth_decimation(const HgEnvBuffers* B, const HgEnvParams* P, const float* dof_frames, int32_t decimation,
                                           const float* root_frame, const float* contact_frame, const float* rigid_frame,
                                           int64_t N, void* stream) {
    HG_REQUIRE(B); HG_REQUIRE(P); HG_REQUIRE(B->actions); HG_REQUIRE(B->dof_state); HG_REQUIRE(B->torques);
    HG_REQUIRE(B->root_states); HG_REQUIRE(B->contact_forces); HG_REQUIRE(B->rigid_state);
    HG_REQUIRE(dof_frames); HG_REQUIRE(root_frame); HG_REQUIRE(contact_frame); HG_REQUIRE(rigid_frame);
    if (N <= 0 || N > (1 << 22) || decimation < 1) return hg_fail(HG_E_SIZE, "hg_env_synth_decimation: bad N / decimation");
    if (!hg_aligned16(B->dof_state) || (reinterpret_cast<uintptr_t>(dof_frames) & 7u)) return hg_fail(HG_E_ALIGN, "hg_env_synth_decimation: dof tensors must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t work = N * P->num_bodies * 13 / 4;
    int grid = (int)((work + 255) / 256);
    if (grid > 4 * HG_NUM_SMS) grid = 4 * HG_NUM_SMS;
    if (grid < 1) grid = 1;
    synth_decimation_kernel<<<grid, 256, 0, st>>>(*P, *B, reinterpret_cast<const float2*>(dof_frames), decimation, root_frame, contact_frame,
                                                  rigid_frame, (int)N, P->num_bodies);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_env_synth_decimation");
}

extern "C" int32_t hg_env_post_physics(const HgEnvBuffers* B, const HgEnvParams* P, const HgEnvNoise* Z,
                                       uint32_t phases, int64_t common_step_counter, int64_t N, void* stream) {
    if (int32_t rc = check_buffers(B)) return rc;
    HG_REQUIRE(P); HG_REQUIRE(Z);
    if (N <= 0 || N > (1 << 26)) return hg_fail(HG_E_SIZE, "hg_env_post_physics: bad N");
    if ((phases & HG_PHASE_OBS) && (B->obs_out == B->obs_buf || B->priv_out == B->privileged_obs_buf))
        return hg_fail(HG_E_ARG, "hg_env_post_physics: obs_out / priv_out must not alias the history inputs");
    if ((B->obs_pitch && B->obs_pitch < 705) || (B->priv_pitch && B->priv_pitch < 219))
        return hg_fail(HG_E_SIZE, "hg_env_post_physics: observation pitch smaller than the row width");
    if (P->num_bodies > HG_MAX_BODIES) return hg_fail(HG_E_SIZE, "hg_env_post_physics: num_bodies > 16");
    if (P->num_bodies <= 0 || P->n_term > HG_MAX_CONTACT_BODIES || P->n_pen > HG_MAX_CONTACT_BODIES)
        return hg_fail(HG_E_ARG, "hg_env_post_physics: bad body indices");
    if ((phases & ~HG_PHASE_STEP_ALL) || phases == 0) return hg_fail(HG_E_ARG, "hg_env_post_physics: bad phase mask");
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(post_physics_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnvSmem));
        cudaFuncSetAttribute(post_physics_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnvSmem));
        cudaFuncSetAttribute(post_physics_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnvSmem));
        attr_set[dev] = true;
    }
    int grid = (int)((N + HG_ENVS_PER_CTA - 1) / HG_ENVS_PER_CTA);
    // one CTA per 32-env tile (the kernel body is a tile loop, so a capped persistent grid also works, but measured no
    // gain on B200).  CTA width by grid depth: 4 role warps + 12 streaming warps with <= 1 tile per SM (the 4096-env training
    // configuration), 4 + 4 warps (2 CTAs per SM) for deeper grids (measured at N = 65536: 172 us vs 186 us for 4 + 0).
    static int env_threads = -1;                          // HG_ENV_CTA=128|256|512 pins the width (profiling)
    if (env_threads < 0) { const char* v = getenv("HG_ENV_CTA"); env_threads = v ? atoi(v) : 0; }
    // (the 8-warp variant measured no better than either neighbour -- 49 us at N=4096, 84 vs 72 us at N=16384 -- so it
    // is only reachable through HG_ENV_CTA)
    const int width = env_threads ? env_threads : (grid <= HG_NUM_SMS ? 512 : 256);
    // HG_ENV_L2_PREFETCH=0|1|2: bulk L2 prefetch of each tile's history rows at tile start (1: grids deeper than one tile per
    // SM only, 2: always)
    static int l2pf_mode = -1;
    if (l2pf_mode < 0) { const char* v = getenv("HG_ENV_L2_PREFETCH"); l2pf_mode = v ? atoi(v) : 0; }
    const int l2pf = l2pf_mode == 2 || (l2pf_mode == 1 && grid > HG_NUM_SMS);
    if (width == 512) post_physics_kernel<512><<<grid, 512, sizeof(EnvSmem), st>>>(*P, *B, *Z, phases, common_step_counter, (int)N, g_env_trace, l2pf);
    else if (width == 256) post_physics_kernel<256><<<grid, 256, sizeof(EnvSmem), st>>>(*P, *B, *Z, phases, common_step_counter, (int)N, g_env_trace, l2pf);
    else post_physics_kernel<128><<<grid, 128, sizeof(EnvSmem), st>>>(*P, *B, *Z, phases, common_step_counter, (int)N, g_env_trace, l2pf);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_env_post_physics");
}
