"""XBot-L task + trainer configuration (values of reference envs/custom/humanoid_config.py:33-261).

These numbers end up, via LeggedRobot._parse_cfg / _native_params, in the HgEnvParams constant block of the
fused env kernel; tests/test_host_api.py compares the whole tree with a dump of the reference's own classes
(tests/golden/cfg_dump.json)."""
from humanoid.envs.base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO

_LEG_JOINTS = ("leg_roll_joint", "leg_yaw_joint", "leg_pitch_joint", "knee_joint", "ankle_pitch_joint", "ankle_roll_joint")


class XBotLCfg(LeggedRobotCfg):
    class env(LeggedRobotCfg.env):
        frame_stack, num_single_obs = 15, 47                        # actor: 15 frames of 47
        c_frame_stack, single_num_privileged_obs = 3, 73            # critic: 3 frames of 73
        num_observations = int(frame_stack * num_single_obs)
        num_privileged_obs = int(c_frame_stack * single_num_privileged_obs)
        num_actions, num_envs, episode_length_s = 12, 4096, 24
        use_ref_actions = False

    class safety:
        pos_limit = vel_limit = 1.0
        torque_limit = 0.85                                          # fraction of the URDF effort limits

    class asset(LeggedRobotCfg.asset):
        name, file = "XBot-L", "{LEGGED_GYM_ROOT_DIR}/resources/robots/XBot/urdf/XBot-L.urdf"
        foot_name, knee_name = "ankle_roll", "knee"
        terminate_after_contacts_on = ["base_link"]
        penalize_contacts_on = ["base_link"]
        self_collisions = 0
        flip_visual_attachments = replace_cylinder_with_capsule = fix_base_link = False

    class terrain(LeggedRobotCfg.terrain):
        mesh_type = "plane"
        curriculum = measure_heights = False
        static_friction = dynamic_friction = 0.6
        restitution = 0.0
        terrain_length = terrain_width = 8.0
        num_rows = num_cols = 20
        max_init_terrain_level = 10
        terrain_proportions = [0.2, 0.2, 0.4, 0.1, 0.1, 0, 0]

    class noise:
        add_noise, noise_level = True, 0.6

        class noise_scales:
            dof_pos, dof_vel, ang_vel, lin_vel, quat, height_measurements = 0.05, 0.5, 0.1, 0.05, 0.03, 0.1

    class init_state(LeggedRobotCfg.init_state):
        pos = [0.0, 0.0, 0.95]
        default_joint_angles = {side + joint: 0.0 for side in ("left_", "right_") for joint in _LEG_JOINTS}

    class control(LeggedRobotCfg.control):
        stiffness = {"leg_roll": 200.0, "leg_pitch": 350.0, "leg_yaw": 200.0, "knee": 350.0, "ankle": 15}
        damping = dict.fromkeys(("leg_roll", "leg_pitch", "leg_yaw", "knee", "ankle"), 10)
        action_scale, decimation = 0.25, 10                         # 100 Hz policy on 1 kHz physics

    class sim(LeggedRobotCfg.sim):
        dt, substeps, up_axis = 0.001, 1, 1

        class physx(LeggedRobotCfg.sim.physx):
            num_threads, solver_type = 10, 1
            num_position_iterations, num_velocity_iterations = 4, 1
            contact_offset, rest_offset = 0.01, 0.0
            bounce_threshold_velocity, max_depenetration_velocity = 0.1, 1.0
            max_gpu_contact_pairs = 2 ** 23
            default_buffer_size_multiplier, contact_collection = 5, 2

    class domain_rand:
        randomize_friction, friction_range = True, [0.1, 2.0]
        randomize_base_mass, added_mass_range = True, [-5.0, 5.0]
        push_robots, push_interval_s = True, 4
        max_push_vel_xy, max_push_ang_vel = 0.2, 0.4
        action_delay, action_noise = 0.5, 0.02

    class commands(LeggedRobotCfg.commands):
        num_commands, resampling_time, heading_command = 4, 8.0, True

        class ranges:
            lin_vel_x = [-0.3, 0.6]
            lin_vel_y, ang_vel_yaw = [-0.3, 0.3], [-0.3, 0.3]
            heading = [-3.14, 3.14]

    class rewards:
        base_height_target, target_feet_height = 0.89, 0.06
        min_dist, max_dist = 0.2, 0.5
        target_joint_pos_scale, cycle_time = 0.17, 0.64
        only_positive_rewards, tracking_sigma, max_contact_force = True, 5, 700

        class scales:                                               # evaluated in alphabetical order (dir())
            action_smoothness, base_acc, base_height, collision = -0.002, 0.2, 0.2, -1.0
            default_joint_pos, dof_acc, dof_vel = 0.5, -1e-7, -5e-4
            feet_air_time, feet_clearance, feet_contact_forces, feet_contact_number, feet_distance = 1.0, 1.0, -0.01, 1.2, 0.2
            foot_slip, joint_pos, knee_distance, low_speed, orientation = -0.05, 1.6, 0.2, 0.2, 1.0
            torques, track_vel_hard, tracking_ang_vel, tracking_lin_vel, vel_mismatch_exp = -1e-5, 0.5, 1.1, 1.2, 0.5

    class normalization:
        clip_observations = clip_actions = 18.0

        class obs_scales:
            lin_vel, ang_vel, dof_pos, dof_vel, quat, height_measurements = 2.0, 1.0, 1.0, 0.05, 1.0, 5.0


class XBotLCfgPPO(LeggedRobotCfgPPO):
    seed, runner_class_name = 5, "OnPolicyRunner"

    class policy:
        init_noise_std = 1.0
        actor_hidden_dims, critic_hidden_dims = [512, 256, 128], [768, 256, 128]

    class algorithm(LeggedRobotCfgPPO.algorithm):
        entropy_coef, learning_rate = 0.001, 1e-5
        num_learning_epochs, num_mini_batches = 2, 4
        gamma, lam = 0.994, 0.9

    class runner:
        policy_class_name, algorithm_class_name = "ActorCritic", "PPO"
        num_steps_per_env, max_iterations, save_interval = 60, 3001, 100
        experiment_name, run_name = "XBot_ppo", ""
        resume, resume_path = False, None
        load_run = checkpoint = -1
