"""IsaacGymPhysics: the real-physics backend behind the PhysicsBackend seam.

Everything the reference does with `isaacgym` lives here, and nothing else does it: sim / ground / terrain creation
(reference envs/base/base_task.py:45, envs/custom/humanoid_env.py:145-163, legged_robot.py:543-586), asset loading and
the per-env actor loop with its friction / base-mass randomisation callbacks (legged_robot.py:257-302,588-681), the
four state tensors (legged_robot.py:438-457), the decimation sub-step (legged_robot.py:96-101) and the state writes
of resets and pushes (legged_robot.py:371-373,395-397, humanoid_env.py:97-98).  The env above it
(envs/base/legged_robot.py of this package) keeps calling the PhysicsBackend methods and launching the fused sm_100a
kernels on the tensors PhysX writes -- with `use_gpu_pipeline` those are CUDA tensors shared with the simulator, so the
hot path stays copy-free.

Isaac Gym Preview 4 ships no sm_100 build (and is not in this image), so on a B200 this module is exercised against the
test-only functional fake in tests/golden/fake_isaacgym (tests/test_isaacgym_adapter.py, tests/test_env_gpu.py); with a
Blackwell-capable build on PYTHONPATH, `HG_PHYSICS=isaacgym` (or the default 'auto') selects it unchanged.  The module
imports `isaacgym` at import time on purpose: select it only when the package is importable.
"""
import os

import numpy as np
import torch
from isaacgym import gymapi, gymtorch          # noqa: E402  (must be importable: see humanoid.physics.make_physics)

from humanoid import LEGGED_GYM_ROOT_DIR
from humanoid.physics import PhysicsBackend

_ASSET_OPTION_FIELDS = ("default_dof_drive_mode", "collapse_fixed_joints", "replace_cylinder_with_capsule",
                        "flip_visual_attachments", "fix_base_link", "density", "angular_damping", "linear_damping",
                        "max_angular_velocity", "max_linear_velocity", "armature", "thickness", "disable_gravity")


def _gym_sim_params(sim_params):
    """The env's SimParams (humanoid.utils.helpers) as a gymapi.SimParams: same fields, copied one by one."""
    p = gymapi.SimParams()
    for k in ("dt", "substeps", "up_axis", "use_gpu_pipeline"):
        if hasattr(sim_params, k):
            setattr(p, k, getattr(sim_params, k))
    g = getattr(sim_params, "gravity", None)
    if g is not None:
        p.gravity = gymapi.Vec3(*g) if isinstance(g, (list, tuple)) else g
    physx = getattr(sim_params, "physx", None)
    for k, v in (vars(physx).items() if physx is not None else ()):
        if not k.startswith("_"):
            setattr(p.physx, k, v)
    return p


class IsaacGymPhysics(PhysicsBackend):
    def __init__(self, num_envs, device, cfg, env_origins, sim_params, physics_engine=None, sim_device_id=0,
                 graphics_device_id=None, custom_origins=False):
        self.num_envs, self.device, self.cfg = num_envs, torch.device(device), cfg
        self.gym = gymapi.acquire_gym()
        self._gym_params = _gym_sim_params(sim_params)
        engine = gymapi.SIM_PHYSX if physics_engine is None else physics_engine
        gfx = sim_device_id if graphics_device_id is None else graphics_device_id
        self.sim = self.gym.create_sim(sim_device_id, gfx, engine, self._gym_params)
        self._terrain_added = False
        self._env_origins = env_origins.detach().cpu()
        self._custom_origins = custom_origins
        self._prepared = False
        self._load_asset()

    # ---- construction ------------------------------------------------------------------------------------------
    def _load_asset(self):                                        # legged_robot.py:597-630
        a = self.cfg.asset
        path = a.file.format(LEGGED_GYM_ROOT_DIR=LEGGED_GYM_ROOT_DIR)
        opts = gymapi.AssetOptions()
        for k in _ASSET_OPTION_FIELDS:
            setattr(opts, k, getattr(a, k))
        gym = self.gym
        self.asset = gym.load_asset(self.sim, os.path.dirname(path), os.path.basename(path), opts)
        self.num_dof = gym.get_asset_dof_count(self.asset)
        self.body_names = list(gym.get_asset_rigid_body_names(self.asset))
        self.dof_names = list(gym.get_asset_dof_names(self.asset))
        self.num_bodies = len(self.body_names)
        self._dof_props = gym.get_asset_dof_properties(self.asset)
        self._shape_props = gym.get_asset_rigid_shape_properties(self.asset)

    def dof_properties(self):
        p = self._dof_props
        return {k: [float(x) for x in p[k]] for k in ("lower", "upper", "velocity", "effort")}

    def add_terrain(self, terrain, mesh_type):                    # legged_robot.py:553-586
        tc = terrain.cfg
        if mesh_type == "heightfield":
            hp = gymapi.HeightFieldParams()
            hp.column_scale = hp.row_scale = tc.horizontal_scale
            hp.vertical_scale = tc.vertical_scale
            hp.nbRows, hp.nbColumns = terrain.tot_cols, terrain.tot_rows
            hp.transform.p.x = hp.transform.p.y = -tc.border_size
            hp.transform.p.z = 0.0
            hp.static_friction, hp.dynamic_friction, hp.restitution = tc.static_friction, tc.dynamic_friction, tc.restitution
            self.gym.add_heightfield(self.sim, terrain.heightsamples, hp)
        elif mesh_type == "trimesh":
            tp = gymapi.TriangleMeshParams()
            tp.nb_vertices, tp.nb_triangles = terrain.vertices.shape[0], terrain.triangles.shape[0]
            tp.transform.p.x = tp.transform.p.y = -tc.border_size
            tp.transform.p.z = 0.0
            tp.static_friction, tp.dynamic_friction, tp.restitution = tc.static_friction, tc.dynamic_friction, tc.restitution
            self.gym.add_triangle_mesh(self.sim, terrain.vertices.flatten(order="C"), terrain.triangles.flatten(order="C"), tp)
        else:
            raise ValueError(f"add_terrain: mesh_type {mesh_type!r}")
        self._terrain_added = True

    def _add_ground_plane(self):                                  # legged_robot.py:543-551
        pp, tc = gymapi.PlaneParams(), self.cfg.terrain
        pp.normal = gymapi.Vec3(0.0, 0.0, 1.0)
        pp.static_friction, pp.dynamic_friction, pp.restitution = tc.static_friction, tc.dynamic_friction, tc.restitution
        self.gym.add_ground(self.sim, pp)

    def create_actors(self):
        """One env + one actor per robot, with the reference's randomisation callbacks (legged_robot.py:257-302,638-665).
        Returns (env_frictions (N,1), body_mass (N,1)) as CPU tensors; LeggedRobot adopts them (critic observations)."""
        gym, cfg, N = self.gym, self.cfg, self.num_envs
        if not self._terrain_added and cfg.terrain.mesh_type == "plane":
            self._add_ground_plane()
        dr = cfg.domain_rand
        env_frictions, body_mass = torch.zeros(N, 1), torch.zeros(N, 1)
        friction_coeffs = None
        if dr.randomize_friction:                                 # 256 buckets, drawn once (legged_robot.py:257-270)
            lo, hi = dr.friction_range
            ids = torch.randint(0, 256, (N, 1))
            friction_coeffs = ((hi - lo) * torch.rand(256, 1) + lo)[ids]
        self.friction_coeffs = friction_coeffs
        pose = gymapi.Transform()
        zero = gymapi.Vec3(0.0, 0.0, 0.0)
        self.envs, self.actor_handles = [], []
        per_row = int(np.sqrt(N))
        for i in range(N):
            env = gym.create_env(self.sim, zero, zero, per_row)
            pos = self._env_origins[i].clone()                    # the first reset_idx() writes the real initial state
            pos[:2] += 2.0 * torch.rand(2) - 1.0                  # xy within 1 m of the origin (:643-645)
            pose.p = gymapi.Vec3(float(pos[0]), float(pos[1]), float(pos[2]))
            if friction_coeffs is not None:
                for sp in self._shape_props:
                    sp.friction = float(friction_coeffs[i])
                env_frictions[i] = friction_coeffs[i]
            gym.set_asset_rigid_shape_properties(self.asset, self._shape_props)
            actor = gym.create_actor(env, self.asset, pose, cfg.asset.name, i, cfg.asset.self_collisions, 0)
            gym.set_actor_dof_properties(env, actor, self._dof_props)
            props = gym.get_actor_rigid_body_properties(env, actor)
            if dr.randomize_base_mass:                            # :296-302
                lo, hi = dr.added_mass_range
                props[0].mass += np.random.uniform(lo, hi)
            body_mass[i] = props[0].mass
            gym.set_actor_rigid_body_properties(env, actor, props, recomputeInertia=True)
            self.envs.append(env)
            self.actor_handles.append(actor)
        return env_frictions, body_mass

    def body_index(self, name):                                   # legged_robot.py:667-681
        return self.gym.find_actor_rigid_body_handle(self.envs[0], self.actor_handles[0], name)

    def prepare(self):
        """gym.prepare_sim + the four state tensors (base_task.py:96, legged_robot.py:438-457)."""
        gym, sim = self.gym, self.sim
        gym.prepare_sim(sim)
        self.root_states = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
        self.dof_state = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim))
        self.contact_forces = gymtorch.wrap_tensor(gym.acquire_net_contact_force_tensor(sim))
        self.rigid_state = gymtorch.wrap_tensor(gym.acquire_rigid_body_state_tensor(sim))
        if self.root_states.device != self.device:
            raise RuntimeError(
                f"Isaac Gym state tensors live on {self.root_states.device}, the env on {self.device}: the fused kernels "
                "need the GPU pipeline (sim_params.use_gpu_pipeline = True with a cuda sim_device)")
        gym.refresh_dof_state_tensor(sim)
        gym.refresh_actor_root_state_tensor(sim)
        gym.refresh_net_contact_force_tensor(sim)
        gym.refresh_rigid_body_state_tensor(sim)
        self._prepared = True

    # ---- stepping (legged_robot.py:96-101,124-126) --------------------------------------------------------------
    def set_dof_actuation_force_tensor(self, torques):
        self.gym.set_dof_actuation_force_tensor(self.sim, gymtorch.unwrap_tensor(torques))

    def simulate(self):
        self.gym.simulate(self.sim)
        if self.device.type == "cpu":
            self.gym.fetch_results(self.sim, True)

    def fetch_results(self):
        self.gym.fetch_results(self.sim, True)

    def refresh_dof_state_tensor(self):
        self.gym.refresh_dof_state_tensor(self.sim)

    def refresh_actor_root_state_tensor(self):
        self.gym.refresh_actor_root_state_tensor(self.sim)

    def refresh_net_contact_force_tensor(self):
        self.gym.refresh_net_contact_force_tensor(self.sim)

    def refresh_rigid_body_state_tensor(self):
        self.gym.refresh_rigid_body_state_tensor(self.sim)

    # ---- state injection (legged_robot.py:371-373,395-397; humanoid_env.py:97-98) --------------------------------
    def set_dof_state_tensor_indexed(self, dof_state, env_ids_int32, count):
        self.gym.set_dof_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(dof_state),
                                              gymtorch.unwrap_tensor(env_ids_int32), count)

    def set_actor_root_state_tensor(self, root_states):
        self.gym.set_actor_root_state_tensor(self.sim, gymtorch.unwrap_tensor(root_states))

    def set_actor_root_state_tensor_indexed(self, root_states, env_ids_int32, count):
        self.gym.set_actor_root_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(root_states),
                                                     gymtorch.unwrap_tensor(env_ids_int32), count)

    def apply_env_writes(self, reset_ids, scratch, pushed):
        """The fused env kernel rewrote root / dof rows in the shared tensors; tell the simulator which.  The indexed
        setters take a host-side count, so this reads the kernel's reset counter back (one 4-byte D2H per env step; the
        reference pays a `nonzero()` sync at the same place, legged_robot.py:139-140)."""
        if pushed:
            self.set_actor_root_state_tensor(self.root_states)
        n = int(scratch[3].item())
        if n:
            ids = reset_ids[:n].contiguous()
            self.set_dof_state_tensor_indexed(self.dof_state, ids, n)
            self.set_actor_root_state_tensor_indexed(self.root_states, ids, n)
