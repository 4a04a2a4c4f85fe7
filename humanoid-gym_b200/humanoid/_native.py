"""ctypes binding of libhg_b200.so (C ABI in include/hg_b200.h).

The product path has NO fallback: if the library is missing or a call fails this
module raises.  PyTorch is used only for device memory and streams -- every
argument that crosses this boundary is a raw pointer, a size or a POD struct.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("HG_B200_LIB", os.path.join(_PKG, "lib", "libhg_b200.so"))

NUM_DOF, NUM_REWARDS, OBS1, PRIV1, OBS_FRAMES, PRIV_FRAMES = 12, 22, 47, 73, 15, 3
MAX_CONTACT_BODIES, MAX_LAYERS = 4, 8

PHASE_COUNTERS, PHASE_CALLBACK, PHASE_TERMINATE, PHASE_REWARD = 0x01, 0x02, 0x04, 0x08
PHASE_RESET, PHASE_OBS, PHASE_LAST, PHASE_STEP_ALL = 0x10, 0x20, 0x40, 0x7F
STEP_FROM_DEVICE = (1 << 64) - 1

f32, i32, i64, u64, u8 = C.c_float, C.c_int32, C.c_int64, C.c_uint64, C.c_uint8
PF = C.c_void_p     # device pointers travel as integers (tensor.data_ptr())


class EnvParams(C.Structure):
    _fields_ = [
        ("dt", f32), ("cycle_time", f32),
        ("clip_actions", f32), ("clip_obs", f32), ("action_scale", f32),
        ("action_delay", f32), ("action_noise", f32),
        ("cmd_x_lo", f32), ("cmd_x_span", f32), ("cmd_y_lo", f32), ("cmd_y_span", f32),
        ("cmd_heading_lo", f32), ("cmd_heading_span", f32),
        ("push_vel_lo", f32), ("push_vel_span", f32), ("push_ang_lo", f32), ("push_ang_span", f32),
        ("dof_reset_lo", f32), ("dof_reset_span", f32),
        ("target_joint_pos_scale", f32), ("target_feet_height", f32), ("base_height_target", f32),
        ("min_dist", f32), ("max_dist", f32), ("max_contact_force", f32), ("tracking_sigma", f32),
        ("obs_scale_lin_vel", f32), ("obs_scale_ang_vel", f32), ("obs_scale_dof_pos", f32),
        ("obs_scale_dof_vel", f32), ("obs_scale_quat", f32),
        ("noise_level", f32), ("max_episode_length_s", f32),
        ("add_noise", i32), ("only_positive_rewards", i32), ("heading_command", i32), ("push_robots", i32),
        ("resample_period", i32), ("push_interval", i32),
        ("max_episode_length", i64),
        ("num_bodies", i32), ("feet", i32 * 2), ("knees", i32 * 2),
        ("n_term", i32), ("term_bodies", i32 * MAX_CONTACT_BODIES),
        ("n_pen", i32), ("pen_bodies", i32 * MAX_CONTACT_BODIES),
        ("reward_scales", f32 * NUM_REWARDS),
        ("p_gains", f32 * NUM_DOF), ("d_gains", f32 * NUM_DOF), ("torque_limits", f32 * NUM_DOF),
        ("default_dof_pos", f32 * NUM_DOF),
        ("noise_scale_vec", f32 * OBS1),
        ("base_init_state", f32 * 13),
    ]


_ENV_BUFFER_NAMES = (
    "root_states", "dof_state", "contact_forces", "rigid_state", "actions", "last_actions",
    "last_last_actions", "torques", "last_dof_vel", "last_root_vel", "commands", "episode_length_buf",
    "reset_buf", "time_out_buf", "extras_time_outs", "base_lin_vel", "base_ang_vel", "projected_gravity",
    "base_euler_xyz", "feet_air_time", "last_contacts", "feet_height", "last_feet_z", "ref_dof_pos",
    "rand_push_force", "rand_push_torque", "env_frictions", "body_mass", "env_origins", "episode_sums",
    "episode_means", "rew_terms", "obs_buf", "privileged_obs_buf", "obs_out", "priv_out", "rew_buf", "reset_ids",
    "scratch")


class EnvBuffers(C.Structure):
    _fields_ = [(n, PF) for n in _ENV_BUFFER_NAMES] + [("obs_pitch", i64), ("priv_pitch", i64)]


class EnvNoise(C.Structure):
    _fields_ = [("u_cmd_cb", PF), ("u_cmd_rs", PF), ("u_dof", PF), ("u_push", PF), ("z_obs", PF),
                ("seed", u64), ("step", u64), ("use_device_counters", i32), ("_pad", i32)]


class MlpDesc(C.Structure):
    _fields_ = [("n_layers", i32), ("dims", i32 * (MAX_LAYERS + 1)),
                ("w_off", i64 * MAX_LAYERS), ("b_off", i64 * MAX_LAYERS), ("ldw", i64 * MAX_LAYERS)]


class Transition(C.Structure):
    _fields_ = [(n, PF) for n in ("obs", "priv_obs", "actions", "rewards", "dones", "time_outs", "values",
                                  "log_prob", "mu", "sigma")] + [("obs_pitch", i64), ("priv_pitch", i64)]


class Storage(C.Structure):
    _fields_ = [(n, PF) for n in ("observations", "privileged_observations", "actions", "rewards", "dones",
                                  "values", "actions_log_prob", "mu", "sigma", "returns", "advantages")] + \
               [("T", i32), ("num_obs", i32), ("num_priv", i32), ("num_actions", i32)]


class Split(C.Structure):
    """x ~= hi + lo as two bf16 planes: element (r, c) of plane k at p[k * plane + r * ld + c] (include/hg_b200.h)."""
    _fields_ = [("p", PF), ("ld", i64), ("plane", i64)]

    @classmethod
    def of(cls, t):
        """t: (2, rows, ld) int16/uint16/bfloat16 CUDA tensor, planes contiguous."""
        assert t.is_cuda and t.dim() == 3 and t.shape[0] == 2 and t.element_size() == 2 and t.stride(2) == 1
        return cls(t.data_ptr(), t.stride(1), t.stride(0))


class MiniBatch(C.Structure):
    _fields_ = [(n, PF) for n in ("obs", "priv_obs", "actions", "values", "advantages", "returns",
                                  "old_log_prob", "old_mu", "old_sigma")] + [("ld_obs", i64), ("ld_priv", i64),
                                                                             ("obs_split", Split), ("priv_split", Split)]


class PpoLossArgs(C.Structure):
    _fields_ = [(n, PF) for n in ("mean", "value", "std", "actions", "target_values", "advantages", "returns",
                                  "old_log_prob", "old_mu", "old_sigma", "d_mean", "d_value", "grad_std",
                                  "scalars")] + \
               [("clip_param", f32), ("value_loss_coef", f32), ("entropy_coef", f32),
                ("use_clipped_value_loss", i32), ("num_actions", i32), ("inv_B", f32)]


class Gemm(C.Structure):
    _fields_ = [("A", PF), ("B", PF), ("C", PF), ("bias", PF), ("H", PF), ("M", i32), ("N", i32), ("K", i32),
                ("lda", i64), ("ldb", i64), ("ldc", i64), ("ldh", i64), ("a_mn_major", i32), ("b_mn_major", i32),
                ("epilogue", i32), ("passes", i32), ("split_k", i32), ("trust_hw_truncation", i32),
                ("B_lo", PF), ("sample_std", PF), ("sample_eps", PF), ("sample_actions", PF), ("sample_log_prob", PF),
                ("sample_sigma", PF), ("sample_seed", u64), ("sample_step", u64), ("sample_step_dev", PF)]


class MlpFwdOpts(C.Structure):
    _fields_ = [("params_lo", PF), ("std", PF), ("eps", PF), ("actions", PF), ("log_prob", PF), ("sigma", PF),
                ("seed", u64), ("step", u64), ("step_dev", PF)]


class GemmSplit(C.Structure):
    _fields_ = [("A", Split), ("B", Split), ("C", PF), ("ldc", i64), ("Cs", Split), ("bias", PF), ("Hs", Split),
                ("colsum", PF), ("M", i32), ("N", i32), ("K", i32), ("a_mn_major", i32), ("b_mn_major", i32),
                ("epilogue", i32), ("split_k", i32)]


class Terrain(C.Structure):
    """HgTerrain: the rough-terrain height field and curriculum constants (include/hg_b200.h)."""
    _fields_ = [("height_samples", PF), ("rows", i32), ("cols", i32), ("border_size", f32), ("horizontal_scale", f32),
                ("vertical_scale", f32), ("terrain_origins", PF), ("num_levels", i32), ("num_types", i32),
                ("half_env_length", f32), ("max_episode_length_s", f32), ("curriculum", i32), ("_pad", i32)]


_STRUCTS = (EnvParams, EnvBuffers, EnvNoise, MlpDesc, Transition, Storage, MiniBatch, PpoLossArgs, Gemm, Split, GemmSplit, MlpFwdOpts,
            Terrain)


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the humanoid_ppo hot path)")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    sig = {
        "hg_version": (i32, []),
        "hg_last_error": (C.c_char_p, []),
        "hg_launch_count": (i64, []),
        "hg_struct_size": (i64, [i32]),
        "hg_env_pre_physics": (i32, [P(EnvBuffers), P(EnvParams), PF, PF, PF, u64, u64, i64, PF]),
        "hg_env_compute_torques": (i32, [P(EnvBuffers), P(EnvParams), i64, PF]),
        "hg_env_set_trace": (None, [PF]),
        "hg_env_synth_decimation": (i32, [P(EnvBuffers), P(EnvParams), PF, i32, PF, PF, PF, i64, PF]),
        "hg_env_post_physics": (i32, [P(EnvBuffers), P(EnvParams), P(EnvNoise), C.c_uint32, i64, i64, PF]),
        "hg_terrain_get_heights": (i32, [P(Terrain), PF, PF, i32, PF, i64, PF]),
        "hg_terrain_reset_prepare": (i32, [P(Terrain), PF, PF, PF, PF, PF, PF, PF, PF, PF, u64, u64, PF, i64, PF]),
        "hg_terrain_priv_frames": (i32, [PF, i64, i32, PF, PF, i32, f32, f32, PF, PF, PF, i64, i32, i64, PF]),
        "hg_mlp_forward": (i32, [P(MlpDesc), PF, PF, i64, PF, PF, i64, PF]),
        "hg_mlp_forward_ex": (i32, [P(MlpDesc), PF, PF, i64, PF, PF, i64, P(MlpFwdOpts), PF]),
        "hg_tf32_residual": (i32, [PF, PF, i64, PF]),
        "hg_actor_critic_counters_size": (i64, [i64]),
        "hg_actor_critic_set_trace": (None, [PF]),
        "hg_gemm_bf16x3_set_trace": (None, [PF]),
        "hg_actor_critic_forward": (i32, [P(MlpDesc), P(MlpDesc), PF, PF, PF, i64, PF, i64, PF, PF, PF, PF, PF, PF, P(MlpFwdOpts), PF, i64, PF]),
        "hg_f16_weight_scale": (C.c_float, []),
        "hg_split_f16": (i32, [PF, i64, P(Split), i64, i64, C.c_float, PF]),
        "hg_actor_critic_f16_scratch_elems": (i64, [P(MlpDesc), P(MlpDesc), i64]),
        "hg_actor_critic_forward_f16": (i32, [P(MlpDesc), P(MlpDesc), PF, PF, i64, PF, i64, PF, i64, PF, PF, PF, P(MlpFwdOpts), PF, i64, PF]),
        "hg_mlp_backward": (i32, [P(MlpDesc), PF, PF, i64, PF, PF, PF, PF, i64, PF]),
        "hg_gemm_tf32": (i32, [P(Gemm), PF]),
        "hg_set_gemm_mode": (i32, [i32]),
        "hg_split_bf16": (i32, [PF, i64, P(Split), i64, i64, PF]),
        "hg_unsplit_bf16": (i32, [P(Split), PF, i64, i64, i64, PF]),
        "hg_gemm_bf16x3": (i32, [P(GemmSplit), PF]),
        "hg_mlp_forward_split": (i32, [P(MlpDesc), PF, PF, i64, P(Split), PF, PF, i64, PF]),
        "hg_mlp_backward_split": (i32, [P(MlpDesc), PF, PF, i64, P(Split), PF, PF, PF, PF, i64, PF]),
        "hg_policy_sample": (i32, [PF, PF, PF, u64, u64, PF, PF, PF, PF, i64, i32, PF]),
        "hg_storage_add": (i32, [P(Storage), P(Transition), i32, f32, i64, PF]),
        "hg_set_gae_mode": (i32, [i32]),
        "hg_gae": (i32, [P(Storage), PF, f32, f32, PF, i32, i64, PF]),
        "hg_adv_normalise": (i32, [P(Storage), PF, i64, PF]),
        "hg_minibatch_gather": (i32, [P(Storage), PF, P(MiniBatch), i64, PF]),
        "hg_ppo_loss_fwd_bwd": (i32, [P(PpoLossArgs), i64, PF]),
        "hg_grad_sqnorm": (i32, [PF, i64, PF, PF]),
        "hg_clip_adam_step": (i32, [PF, PF, PF, PF, PF, f32, PF, PF, f32, f32, f32, f32, i64, PF]),
        "hg_clip_adam_step_stats": (i32, [PF, PF, PF, PF, PF, f32, PF, PF, f32, f32, f32, f32, i64, PF, PF, i32, PF]),
        "hg_randperm": (i32, [i64, u64, u64, PF, PF]),
        "hg_adapt_lr": (i32, [PF, C.c_double, PF, PF]),
        "hg_episode_book_step": (i32, [PF, PF, PF, PF, PF, PF, PF, PF, i32, i64, PF]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here == the library does not export the ABI
        fn.restype, fn.argtypes = res, args
    for k, st in enumerate(_STRUCTS):
        got = lib.hg_struct_size(k)
        if got != C.sizeof(st):
            raise NativeError(f"ABI mismatch for {st.__name__}: library {got} B, binding {C.sizeof(st)} B")
    return lib, tuple(sig)


lib, EXPORTS = _load()


def check(rc, what=""):
    if rc != 0:
        msg = lib.hg_last_error().decode(errors="replace")
        raise NativeError(f"{what or 'libhg_b200'} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "native kernels need contiguous CUDA tensors"
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def launch_count():
    return int(lib.hg_launch_count())
