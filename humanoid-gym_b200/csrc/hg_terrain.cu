// Rough-terrain side of the env step: height sampling around every robot, the terrain curriculum
// of the envs that terminated, and the height-augmented critic frames.  Compiled with -fmad=false like hg_env.cu: each
// function mirrors a chain of separately-rounded fp32 torch ops of the reference (file:line relative to
// /root/reference/humanoid/).  All three kernels are HBM / gather bound and tiny next to the fused env kernel; they only
// run when cfg.terrain.mesh_type is 'heightfield' or 'trimesh' (XBotLCfg ships 'plane').
#include "hg_common.cuh"

namespace {

// LeggedRobot._get_heights, legged_robot.py:759-795; quat_apply_yaw utils/math.py:38-43 with Isaac Gym's
// normalize (x / |x|.clamp(min=1e-9)) and quat_apply (b + w t + xyz x t, t = 2 xyz x b), xyzw quaternions.
// One WARP per env, 8 envs per CTA, no block-level synchronisation: every lane reads the root row (a broadcast), normalises
// the yaw quaternion itself (1 sqrt + 2 divisions per lane, not per point) and owns points lane, lane + 32, ...; per point
// ~10 flops, two IEEE divisions, three int16 gathers, one coalesced store.  What bounds it is the L1 line rate of those
// gathers (a warp's 32 points touch ~16 different 128-byte lines of the height field per load; measured variants: one
// thread per point 80 us, one CTA per env 76 us, a shared-memory copy of the 32 x 32-cell window around the robot 141 us
// at N = 65536 -- the extra barriers cost more than the gathers they saved).
constexpr int HG_HEIGHT_WARPS = 8;
__global__ void __launch_bounds__(HG_HEIGHT_WARPS * 32) get_heights_kernel(HgTerrain T, const float* __restrict__ root, const float* __restrict__ pts,
                                                                          int P, float* __restrict__ heights, int64_t N) {
    const int64_t e = (int64_t)blockIdx.x * HG_HEIGHT_WARPS + (threadIdx.x >> 5);
    if (e >= N) return;
    const int lane = threadIdx.x & 31;
    const float* r = root + e * 13;
    float qz = r[5], qw = r[6];
    float n = sqrtf(qz * qz + qw * qw);              // quat_yaw = (0, 0, z, w): the two zeroed components add nothing
    n = fmaxf(n, 1e-9f);
    qz = qz / n; qw = qw / n;
    const float rx = r[0], ry = r[1];
    for (int p = lane; p < P; p += 32) {
        const float bx = pts[2 * p], by = pts[2 * p + 1];                 // height_points[..., 2] = 0
        const float tx = (0.0f - qz * by) * 2.0f, ty = (qz * bx) * 2.0f;   // t = cross((0,0,z), b) * 2
        const float cx = 0.0f - qz * ty, cy = qz * tx;                     // cross((0,0,z), t)
        float x = (bx + qw * tx) + cx, y = (by + qw * ty) + cy;
        x = x + rx; y = y + ry;                                           // + root_states[:, :3]
        x = x + T.border_size; y = y + T.border_size;                     // points += border_size
        // (points / horizontal_scale).long(): truncation; clamping the float first keeps the conversion defined off the map
        const float fx = fminf(fmaxf(x / T.horizontal_scale, -1.0f), 2147483520.0f), fy = fminf(fmaxf(y / T.horizontal_scale, -1.0f), 2147483520.0f);
        int px = (int)fx, py = (int)fy;
        px = px < 0 ? 0 : (px > T.rows - 2 ? T.rows - 2 : px);            // clip to [0, shape-2]
        py = py < 0 ? 0 : (py > T.cols - 2 ? T.cols - 2 : py);
        const int16_t* hs = T.height_samples + (size_t)px * T.cols + py;
        const int h = min(min((int)hs[0], (int)hs[T.cols]), (int)hs[1]);   // (px,py), (px+1,py), (px,py+1)
        heights[e * P + p] = (float)h * T.vertical_scale;
    }
}

// _update_terrain_curriculum legged_robot.py:400-420 followed by the custom-origin spawn of _reset_root_states :381-384
__global__ void reset_prepare_kernel(HgTerrain T, const uint8_t* __restrict__ reset_buf, const float* __restrict__ root,
                                     const float* __restrict__ commands, int64_t* __restrict__ levels,
                                     const int64_t* __restrict__ types, float* __restrict__ env_origins,
                                     float* __restrict__ spawn, const int64_t* __restrict__ r_level,
                                     const float* __restrict__ u_root, uint64_t seed, uint64_t step,
                                     const uint64_t* __restrict__ step_dev, int N) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N || !reset_buf[e]) return;
    if (step_dev) step = *step_dev;
    float* org = env_origins + (size_t)e * 3;
    if (T.curriculum) {
        const float dx = root[(size_t)e * 13] - org[0], dy = root[(size_t)e * 13 + 1] - org[1];
        const float distance = sqrtf(dx * dx + dy * dy);                       // torch.norm(..., dim=1)
        const bool up = distance > T.half_env_length;
        const float cx = commands[(size_t)e * 4], cy = commands[(size_t)e * 4 + 1];
        const float need = sqrtf(cx * cx + cy * cy) * T.max_episode_length_s * 0.5f;
        const bool down = (distance < need) && !up;
        long long lvl = levels[e] + (up ? 1 : 0) - (down ? 1 : 0);
        if (lvl >= T.num_levels) {                                             // solved the last level: random level
            if (r_level) lvl = r_level[e];
            else {
                HgPhilox r = hg_philox(seed, (uint32_t)e, (uint32_t)step, HG_RNG_LEVEL | ((uint32_t)(step >> 32) << 8), 0);
                lvl = min((int)(hg_u01(r.c[0]) * (float)T.num_levels), T.num_levels - 1);
            }
        } else if (lvl < 0) lvl = 0;
        levels[e] = lvl;
        const float* o = T.terrain_origins + ((size_t)lvl * T.num_types + types[e]) * 3;
        org[0] = o[0]; org[1] = o[1]; org[2] = o[2];
    }
    float u0, u1;
    if (u_root) { u0 = u_root[(size_t)e * 2]; u1 = u_root[(size_t)e * 2 + 1]; }
    else {
        HgPhilox r = hg_philox(seed, (uint32_t)e, (uint32_t)step, HG_RNG_ROOT | ((uint32_t)(step >> 32) << 8), 0);
        u0 = hg_u01(r.c[0]); u1 = hg_u01(r.c[1]);
    }
    // torch_rand_float(-1, 1): (upper - lower) * u + lower
    spawn[(size_t)e * 3] = org[0] + (2.0f * u0 + -1.0f);
    spawn[(size_t)e * 3 + 1] = org[1] + (2.0f * u1 + -1.0f);
    spawn[(size_t)e * 3 + 2] = org[2];
}

// humanoid_env.py:246-248 (frame), :253-258 (append), :264-269 (reset zeroing), legged_robot.py:104-108 (clip).
// grid = (envs, column chunks): no per-element index division; the history shift out[0 : (F-1) W] = in[W : F W] moves as
// 16-byte vectors when W and the pitch are multiples of 4 floats (892 and 2688 are); the new frame is assembled scalar.
template <bool VEC>
__global__ void __launch_bounds__(256) priv_frames_kernel(const float* __restrict__ obs_prev, int64_t obs_pitch, int num_obs,
                                                          const float* __restrict__ root, const float* __restrict__ heights, int P,
                                                          float height_scale, float clip_obs, const uint8_t* __restrict__ reset_buf,
                                                          const float* __restrict__ priv_in, float* __restrict__ priv_out,
                                                          int64_t priv_pitch, int frames) {
    const int64_t e = blockIdx.x;
    const int W = num_obs + P, keep = (frames - 1) * W;
    const bool rz = reset_buf && reset_buf[e];
    const float* in = priv_in + e * priv_pitch + W;
    float* out = priv_out + e * priv_pitch;
    const int t0 = blockIdx.y * blockDim.x + threadIdx.x, stride = gridDim.y * blockDim.x;
    if (VEC) {
        const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int c = t0; c < keep / 4; c += stride)
            reinterpret_cast<float4*>(out)[c] = rz ? z4 : reinterpret_cast<const float4*>(in)[c];
    } else {
        for (int c = t0; c < keep; c += stride) out[c] = rz ? 0.0f : in[c];
    }
    const float rz_m = root[e * 13 + 2] - 0.5f;
    for (int k = t0; k < W; k += stride) {
        float v;
        if (k < num_obs) v = obs_prev[e * obs_pitch + k];
        else v = fminf(fmaxf(rz_m - heights[e * P + (k - num_obs)], -1.0f), 1.0f) * height_scale;
        out[keep + k] = fminf(fmaxf(v, -clip_obs), clip_obs);
    }
}

int32_t check_terrain(const HgTerrain* T, bool need_origins) {
    HG_REQUIRE(T);
    HG_REQUIRE(T->height_samples);
    if (T->rows < 2 || T->cols < 2) return hg_fail(HG_E_SIZE, "HgTerrain: height field smaller than 2x2");
    if (!(T->horizontal_scale > 0.0f)) return hg_fail(HG_E_ARG, "HgTerrain: horizontal_scale must be positive");
    if (need_origins) {
        HG_REQUIRE(T->terrain_origins);
        if (T->num_levels <= 0 || T->num_types <= 0) return hg_fail(HG_E_SIZE, "HgTerrain: bad num_levels / num_types");
    }
    return 0;
}

}  // namespace

extern "C" int32_t hg_terrain_get_heights(const HgTerrain* T, const float* root_states, const float* points_xy, int32_t P,
                                          float* heights, int64_t N, void* stream) {
    if (int32_t rc = check_terrain(T, false)) return rc;
    HG_REQUIRE(root_states); HG_REQUIRE(points_xy); HG_REQUIRE(heights);
    if (N <= 0 || N > (1 << 26) || P <= 0 || P > 4096) return hg_fail(HG_E_SIZE, "hg_terrain_get_heights: bad N or P");
    get_heights_kernel<<<(unsigned)((N + HG_HEIGHT_WARPS - 1) / HG_HEIGHT_WARPS), HG_HEIGHT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        *T, root_states, points_xy, P, heights, N);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_terrain_get_heights");
}

extern "C" int32_t hg_terrain_reset_prepare(const HgTerrain* T, const uint8_t* reset_buf, const float* root_states,
                                            const float* commands, int64_t* terrain_levels, const int64_t* terrain_types,
                                            float* env_origins, float* spawn, const int64_t* r_level, const float* u_root,
                                            uint64_t seed, uint64_t step, const uint64_t* step_dev, int64_t N, void* stream) {
    if (int32_t rc = check_terrain(T, true)) return rc;
    HG_REQUIRE(reset_buf); HG_REQUIRE(root_states); HG_REQUIRE(commands); HG_REQUIRE(terrain_levels); HG_REQUIRE(terrain_types);
    HG_REQUIRE(env_origins); HG_REQUIRE(spawn);
    if (env_origins == spawn) return hg_fail(HG_E_ARG, "hg_terrain_reset_prepare: spawn must not alias env_origins");
    if (N <= 0 || N > (1 << 26)) return hg_fail(HG_E_SIZE, "hg_terrain_reset_prepare: bad N");
    reset_prepare_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        *T, reset_buf, root_states, commands, terrain_levels, terrain_types, env_origins, spawn, r_level, u_root, seed, step,
        step_dev, (int)N);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_terrain_reset_prepare");
}

extern "C" int32_t hg_terrain_priv_frames(const float* obs_prev, int64_t obs_pitch, int32_t num_obs, const float* root_states,
                                          const float* heights, int32_t P, float height_scale, float clip_obs,
                                          const uint8_t* reset_buf, const float* priv_in, float* priv_out, int64_t priv_pitch,
                                          int32_t frames, int64_t N, void* stream) {
    HG_REQUIRE(obs_prev); HG_REQUIRE(root_states); HG_REQUIRE(heights); HG_REQUIRE(priv_in); HG_REQUIRE(priv_out);
    if (priv_in == priv_out) return hg_fail(HG_E_ARG, "hg_terrain_priv_frames: priv_out must not alias priv_in (ping-pong pair)");
    if (N <= 0 || N > (1 << 26) || P <= 0 || num_obs <= 0 || frames <= 0) return hg_fail(HG_E_SIZE, "hg_terrain_priv_frames: bad sizes");
    if (obs_pitch < num_obs || priv_pitch < (int64_t)frames * (num_obs + P)) return hg_fail(HG_E_SIZE, "hg_terrain_priv_frames: pitch smaller than the row");
    const int W = num_obs + P;
    const bool vec = (W % 4 == 0) && (priv_pitch % 4 == 0) && hg_aligned16(priv_in) && hg_aligned16(priv_out);
    const dim3 grid((unsigned)N, 2);
    if (vec)
        priv_frames_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(obs_prev, obs_pitch, num_obs, root_states, heights, P, height_scale,
                                                                         clip_obs, reset_buf, priv_in, priv_out, priv_pitch, frames);
    else
        priv_frames_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(obs_prev, obs_pitch, num_obs, root_states, heights, P, height_scale,
                                                                          clip_obs, reset_buf, priv_in, priv_out, priv_pitch, frames);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_terrain_priv_frames");
}
