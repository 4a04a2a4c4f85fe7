"""Default configuration of a legged robot task and of its PPO trainer.

Same attribute tree and default values as reference envs/base/legged_robot_config.py:34-236 (the config *is* the
API: user code and class_to_dict() address these names); tests/test_host_api.py checks the tree against a dump of
the reference's classes.  Only the XBot-L overrides (envs/custom/humanoid_config.py) reach the kernels."""
from .base_config import BaseConfig


def _grid(lo, hi, n):
    """n evenly spaced points lo .. hi rounded to one decimal (the height-scan sample offsets)."""
    return [round(lo + (hi - lo) * i / (n - 1), 1) for i in range(n)]


class LeggedRobotCfg(BaseConfig):
    class env:
        num_envs, num_actions = 4096, 12
        num_observations, num_privileged_obs = 235, None          # not None -> step() also returns critic observations
        env_spacing = 3.0                                           # grid pitch of env origins on flat ground [m]
        send_timeouts = True                                        # expose time-outs to the algorithm (bootstrapping)
        episode_length_s = 20

    class terrain:
        mesh_type = "trimesh"                                       # none | plane | heightfield | trimesh
        horizontal_scale, vertical_scale, border_size = 0.1, 0.005, 25
        curriculum = measure_heights = True
        static_friction = dynamic_friction = 1.0
        restitution = 0.0
        measured_points_x, measured_points_y = _grid(-0.8, 0.8, 17), _grid(-0.5, 0.5, 11)
        selected, terrain_kwargs = False, None
        max_init_terrain_level = 5
        terrain_length = terrain_width = 8.0
        num_rows, num_cols = 10, 20
        terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2]
        slope_treshold = 0.75

    class commands:
        curriculum, max_curriculum = False, 1.0
        num_commands = 4                                            # lin_vel_x, lin_vel_y, ang_vel_yaw, heading
        resampling_time, heading_command = 10.0, True

        class ranges:
            lin_vel_x, lin_vel_y = [-1.0, 1.0], [-1.0, 1.0]
            ang_vel_yaw, heading = [-1, 1], [-3.14, 3.14]

    class init_state:
        pos, rot = [0.0, 0.0, 1.0], [0.0, 0.0, 0.0, 1.0]            # quaternion xyzw
        lin_vel, ang_vel = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
        default_joint_angles = {"joint_a": 0.0, "joint_b": 0.0}

    class control:
        stiffness, damping = {"joint_a": 10.0, "joint_b": 15.0}, {"joint_a": 1.0, "joint_b": 1.5}
        action_scale, decimation = 0.5, 4

    class asset:
        file, name, foot_name = "", "legged_robot", "None"
        penalize_contacts_on, terminate_after_contacts_on = [], []
        disable_gravity = fix_base_link = False
        collapse_fixed_joints = replace_cylinder_with_capsule = flip_visual_attachments = True
        default_dof_drive_mode, self_collisions = 3, 0
        density, thickness = 0.001, 0.01
        angular_damping = linear_damping = armature = 0.0
        max_angular_velocity = max_linear_velocity = 1000.0

    class domain_rand:
        randomize_friction, friction_range = True, [0.5, 1.25]
        randomize_base_mass, added_mass_range = False, [-1.0, 1.0]
        push_robots, push_interval_s, max_push_vel_xy = True, 15, 1.0

    class rewards:
        only_positive_rewards, tracking_sigma, max_contact_force = True, 0.25, 100.0

        class scales:
            tracking_lin_vel, tracking_ang_vel, feet_air_time = 1.0, 0.5, 1.0
            lin_vel_z, ang_vel_xy, torques, dof_acc, collision = -2.0, -0.05, -0.00001, -2.5e-7, -1.0
            termination = orientation = dof_vel = base_height = feet_stumble = action_rate = stand_still = -0.0

    class normalization:
        clip_observations = clip_actions = 100.0

        class obs_scales:
            lin_vel, ang_vel, dof_pos, dof_vel, height_measurements = 2.0, 0.25, 1.0, 0.05, 5.0

    class noise:
        add_noise, noise_level = True, 1.0

        class noise_scales:
            dof_pos, dof_vel, lin_vel, ang_vel, gravity, height_measurements = 0.01, 1.5, 0.1, 0.2, 0.05, 0.1

    class viewer:
        ref_env, pos, lookat = 0, [10, 0, 6], [11.0, 5, 3.0]

    class sim:
        dt, substeps = 0.005, 1
        gravity, up_axis = [0.0, 0.0, -9.81], 1                     # up_axis 0: y, 1: z

        class physx:
            num_threads, solver_type = 10, 1
            num_position_iterations, num_velocity_iterations = 4, 0
            contact_offset, rest_offset = 0.01, 0.0
            bounce_threshold_velocity, max_depenetration_velocity = 0.5, 1.0
            max_gpu_contact_pairs = 2 ** 23
            default_buffer_size_multiplier, contact_collection = 5, 2


class LeggedRobotCfgPPO(BaseConfig):
    seed, runner_class_name = 1, "OnPolicyRunner"

    class policy:
        init_noise_std = 1.0
        actor_hidden_dims, critic_hidden_dims = [512, 256, 128], [512, 256, 128]

    class algorithm:
        value_loss_coef, use_clipped_value_loss, clip_param = 1.0, True, 0.2
        entropy_coef, learning_rate, schedule, desired_kl = 0.01, 1.0e-3, "adaptive", 0.01
        num_learning_epochs, num_mini_batches = 5, 4
        gamma, lam, max_grad_norm = 0.99, 0.95, 1.0

    class runner:
        policy_class_name, algorithm_class_name = "ActorCritic", "PPO"
        num_steps_per_env, max_iterations, save_interval = 24, 1500, 100
        experiment_name, run_name = "test", ""
        resume, resume_path = False, None
        load_run = checkpoint = -1
