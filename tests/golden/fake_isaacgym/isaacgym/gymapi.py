"""Functional fake of the isaacgym.gymapi surface the reference touches
(list in SURVEY.md section 8c).  TEST-ONLY.  simulate() writes seeded synthetic
state tensors; nothing here is physics.

Two modes (env HG_FAKE_GYM):
  golden (default) -- CPU, simulate() draws fresh random state every sub-step (tests/golden/make_golden.py).
  ring             -- bench reference arm: state tensors live on the sim device (cuda:i with the GPU pipeline),
                      a ring of frames is pre-generated in prepare_sim() by the SAME standalone generator the
                      product's SyntheticPhysics uses (humanoid-gym_b200/humanoid/synthetic_frames.py, loaded by
                      file path -- nothing of the product package is imported), simulate() only advances a
                      counter and refresh_*() copy the current frame: identical per-step "physics" cost in both arms.
"""
import importlib.util
import math
import os
import types
import xml.etree.ElementTree as ET

import numpy as np
import torch

SIM_PHYSX = 1
SIM_FLEX = 0
KEY_ESCAPE = 0
KEY_V = 1


class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = x, y, z, w


class Transform:
    def __init__(self, p=None, r=None):
        self.p = p if p is not None else Vec3()
        self.r = r if r is not None else Quat()


class PlaneParams:
    pass


class AssetOptions:
    pass


class HeightFieldParams:
    def __init__(self):
        self.transform = Transform()


class TriangleMeshParams:
    def __init__(self):
        self.transform = Transform()


class CameraProperties:
    pass


class _PhysX:
    use_gpu = False
    num_subscenes = 0
    num_threads = 0


class SimParams:
    def __init__(self):
        self.dt = float(np.float32(1.0 / 60.0))
        self.substeps = 2
        self.up_axis = 1
        self.use_gpu_pipeline = False
        self.gravity = Vec3(0.0, 0.0, -9.81)
        self.physx = _PhysX()


class _ShapeProps:
    friction = 1.0


class _BodyProps:
    def __init__(self, mass):
        self.mass = mass


# nominal link masses are irrelevant to the hot path except body 0 (base_link)
_BASE_MASS = 5.0


_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))


def _load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(_REPO, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _TableAsset:
    """XBot-L as Isaac Gym reports it, from the product's constant table (used when the URDF did not travel:
    `pip install --target baseline/_ref` copies the python packages only, not resources/)."""

    def __init__(self):
        m = _load_by_path("_hg_xbot_l_model", "humanoid-gym_b200/humanoid/envs/custom/xbot_l_model.py")
        self.body_names = list(m.BODY_NAMES)
        self.dof_names = list(m.DOF_NAMES)
        rows = list(zip(m.DOF_LOWER, m.DOF_UPPER, m.DOF_VELOCITY, m.DOF_EFFORT))
        self.dof_props = np.array(rows, dtype=[("lower", "f4"), ("upper", "f4"), ("velocity", "f4"), ("effort", "f4")])


class _Asset:
    def __init__(self, path, collapse_fixed):
        root = ET.parse(path).getroot()
        joints = root.findall("joint")
        child_of_fixed = set()
        if collapse_fixed:
            for j in joints:
                if j.get("type") == "fixed" and not j.get("dont_collapse"):
                    child_of_fixed.add(j.find("child").get("link"))
        self.body_names = [l.get("name") for l in root.findall("link") if l.get("name") not in child_of_fixed]
        self.dof_names = []
        rows = []
        for j in joints:
            if j.get("type") == "fixed":
                continue
            lim = j.find("limit")
            self.dof_names.append(j.get("name"))
            rows.append((float(lim.get("lower")), float(lim.get("upper")),
                         float(lim.get("velocity")), float(lim.get("effort"))))
        self.dof_props = np.array(rows, dtype=[("lower", "f4"), ("upper", "f4"), ("velocity", "f4"), ("effort", "f4")])


class _Sim:
    def __init__(self, params, device):
        self.params = params
        self.device = device
        self.envs = []
        self.asset = None
        self.tensors = None
        self.step_count = 0
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(1234)
        self.mode = os.environ.get("HG_FAKE_GYM", "golden")
        self.substep = 0
        self.origins = []
        # knobs for the synthetic writer (set by the golden script)
        self.p_base_contact = 0.02
        self.p_contact_flip = 0.15


class Gym:
    """One instance per acquire_gym() call."""

    # ---- construction ----
    def create_sim(self, compute_device, graphics_device, physics_engine, params):
        dev = "cpu"
        if os.environ.get("HG_FAKE_GYM", "golden") == "ring" and getattr(params, "use_gpu_pipeline", False):
            dev = f"cuda:{int(compute_device)}"
        return _Sim(params, dev)

    def add_ground(self, sim, plane_params):
        pass

    def add_heightfield(self, sim, samples, params):
        sim.terrain = ("heightfield", samples.shape, params)

    def add_triangle_mesh(self, sim, vertices, triangles, params):
        assert vertices.size == 3 * params.nb_vertices and triangles.size == 3 * params.nb_triangles
        sim.terrain = ("trimesh", vertices.shape, params)

    def load_asset(self, sim, root, file, options):
        path = os.path.join(root, file)
        if os.path.exists(path):
            sim.asset = _Asset(path, getattr(options, "collapse_fixed_joints", True))
        else:
            sim.asset = _TableAsset()
        return sim.asset

    def get_asset_dof_count(self, a):
        return len(a.dof_names)

    def get_asset_rigid_body_count(self, a):
        return len(a.body_names)

    def get_asset_dof_properties(self, a):
        return a.dof_props.copy()

    def get_asset_rigid_shape_properties(self, a):
        return [_ShapeProps() for _ in range(3)]

    def get_asset_rigid_body_names(self, a):
        return list(a.body_names)

    def get_asset_dof_names(self, a):
        return list(a.dof_names)

    def create_env(self, sim, lower, upper, per_row):
        h = types.SimpleNamespace(idx=len(sim.envs), sim=sim)
        sim.envs.append(h)
        return h

    def set_asset_rigid_shape_properties(self, a, props):
        pass

    def create_actor(self, env, asset, pose, name, group, filt, seg):
        env.body_props = [_BodyProps(_BASE_MASS if i == 0 else 1.0) for i in range(len(asset.body_names))]
        env.sim.origins.append((pose.p.x, pose.p.y, pose.p.z))
        return 0

    def set_actor_dof_properties(self, env, actor, props):
        pass

    def get_actor_rigid_body_properties(self, env, actor):
        return env.body_props

    def set_actor_rigid_body_properties(self, env, actor, props, recomputeInertia=True):
        pass

    def find_actor_rigid_body_handle(self, env, actor, name):
        return env.sim.asset.body_names.index(name)

    def prepare_sim(self, sim):
        n = len(sim.envs)
        nb = len(sim.asset.body_names)
        nd = len(sim.asset.dof_names)
        dev = sim.device
        sim.tensors = dict(
            root=torch.zeros(n, 13, device=dev), dof=torch.zeros(n * nd, 2, device=dev),
            contact=torch.zeros(n * nb, 3, device=dev), rigid=torch.zeros(n * nb, 13, device=dev))
        sim.tensors["root"][:, 6] = 1.0
        sim.tensors["root"][:, 2] = 0.95
        if sim.mode == "ring":
            gen = _load_by_path("_hg_synthetic_frames", "humanoid-gym_b200/humanoid/synthetic_frames.py")
            sim.decimation, sim.ring = 10, 6
            origins = torch.tensor(sim.origins, dtype=torch.float32)
            lim = sim.asset.dof_props
            fr = gen.generate_ring(n, dev, {"lin_vel_x": [-0.3, 0.6], "lin_vel_y": [-0.3, 0.3]}, origins,
                                   [float(x) for x in lim["lower"]], [float(x) for x in lim["upper"]], nb,
                                   decimation=sim.decimation, seed=5, ring=sim.ring)
            K = sim.ring
            sim.frames = dict(root=fr["root"], dof=fr["dof"].view(K * sim.decimation, n * nd, 2),
                              contact=fr["contact"].view(K, n * nb, 3), rigid=fr["rigid"].view(K, n * nb, 13))

    def create_camera_sensor(self, env, props):
        return 0

    def create_viewer(self, sim, props):
        return None

    # ---- tensor API ----
    def acquire_actor_root_state_tensor(self, sim):
        return sim.tensors["root"]

    def acquire_dof_state_tensor(self, sim):
        return sim.tensors["dof"]

    def acquire_net_contact_force_tensor(self, sim):
        return sim.tensors["contact"]

    def acquire_rigid_body_state_tensor(self, sim):
        return sim.tensors["rigid"]

    @staticmethod
    def _frame(sim):
        return ((sim.substep - 1) // sim.decimation) % sim.ring if sim.substep > 0 else 0

    def refresh_dof_state_tensor(self, sim):
        if sim.mode == "ring":
            sim.tensors["dof"].copy_(sim.frames["dof"][(sim.substep - 1) % (sim.ring * sim.decimation)], non_blocking=True)

    def refresh_actor_root_state_tensor(self, sim):
        if sim.mode == "ring":
            sim.tensors["root"].copy_(sim.frames["root"][self._frame(sim)], non_blocking=True)

    def refresh_net_contact_force_tensor(self, sim):
        if sim.mode == "ring":
            sim.tensors["contact"].copy_(sim.frames["contact"][self._frame(sim)], non_blocking=True)

    def refresh_rigid_body_state_tensor(self, sim):
        if sim.mode == "ring":
            sim.tensors["rigid"].copy_(sim.frames["rigid"][self._frame(sim)], non_blocking=True)

    def set_dof_actuation_force_tensor(self, sim, t):
        sim.last_torques = t

    def set_dof_state_tensor_indexed(self, sim, t, ids, n):
        pass

    def set_actor_root_state_tensor(self, sim, t):
        pass

    def set_actor_root_state_tensor_indexed(self, sim, t, ids, n):
        pass

    def fetch_results(self, sim, wait):
        pass

    def simulate(self, sim):
        """Synthetic state writer (in place, like PhysX writing its GPU buffers)."""
        if sim.mode == "ring":
            sim.substep += 1
            return
        sim.step_count += 1
        g = sim.gen
        T = sim.tensors
        n = T["root"].shape[0]
        nb = T["contact"].shape[0] // n
        nd = T["dof"].shape[0] // n
        lim = sim.asset.dof_props

        def randn(*s):
            return torch.randn(*s, generator=g)

        def rand(*s):
            return torch.rand(*s, generator=g)

        root = T["root"]
        root[:, 0:2] += 0.01 * randn(n, 2)
        root[:, 2] = 0.95 + 0.02 * randn(n)
        rpy = 0.1 * randn(n, 3)
        rpy[:, 2] = (2 * rand(n) - 1) * math.pi
        cr, sr = torch.cos(rpy[:, 0] / 2), torch.sin(rpy[:, 0] / 2)
        cp, sp = torch.cos(rpy[:, 1] / 2), torch.sin(rpy[:, 1] / 2)
        cy, sy = torch.cos(rpy[:, 2] / 2), torch.sin(rpy[:, 2] / 2)
        root[:, 3] = sr * cp * cy - cr * sp * sy
        root[:, 4] = cr * sp * cy + sr * cp * sy
        root[:, 5] = cr * cp * sy - sr * sp * cy
        root[:, 6] = cr * cp * cy + sr * sp * sy
        root[:, 7:10] = 0.3 * randn(n, 3)
        root[:, 10:13] = 0.3 * randn(n, 3)

        dof = T["dof"].view(n, nd, 2)
        q = 0.2 * randn(n, nd)
        dof[..., 0] = torch.max(torch.min(q, torch.from_numpy(lim["upper"].copy())), torch.from_numpy(lim["lower"].copy()))
        dof[..., 1] = randn(n, nd)

        contact = T["contact"].view(n, nb, 3)
        contact.zero_()
        feet = [6, 12]
        clock = math.sin(2 * math.pi * sim.step_count * 0.001 / 0.064)
        stance = torch.tensor([clock >= 0, clock < 0]).repeat(n, 1)
        flip = rand(n, 2) < sim.p_contact_flip
        in_contact = stance ^ flip
        fz = (200 + 400 * rand(n, 2)) * in_contact
        # a few values around the 5 N contact threshold and the 700 N penalty knee
        fz = torch.where(rand(n, 2) < 0.05, 10 * rand(n, 2), fz)
        fz = torch.where(rand(n, 2) < 0.05, 600 + 400 * rand(n, 2), fz)
        for k, b in enumerate(feet):
            contact[:, b, 2] = fz[:, k]
            contact[:, b, 0:2] = 20 * randn(n, 2)
        base_hit = rand(n) < sim.p_base_contact
        contact[:, 0, :] = base_hit.unsqueeze(1) * (2.0 + 5 * rand(n, 3))
        small = rand(n) < 0.05
        contact[:, 0, :] += (small & ~base_hit).unsqueeze(1) * 0.2 * rand(n, 3)

        rigid = T["rigid"].view(n, nb, 13)
        rigid.zero_()
        swing = (~in_contact).float()
        for k, b in enumerate(feet):
            side = 0.15 if k == 0 else -0.15
            rigid[:, b, 0] = root[:, 0] + 0.05 * randn(n)
            rigid[:, b, 1] = root[:, 1] + side + 0.08 * randn(n)
            rigid[:, b, 2] = 0.05 + 0.06 * swing[:, k] * abs(clock) + 0.004 * randn(n)
            rigid[:, b, 7:9] = 0.3 * randn(n, 2) * (0.2 + swing[:, k:k + 1])
        for k, b in enumerate([4, 10]):
            side = 0.12 if k == 0 else -0.12
            rigid[:, b, 0] = root[:, 0] + 0.03 * randn(n)
            rigid[:, b, 1] = root[:, 1] + side + 0.05 * randn(n)
            rigid[:, b, 2] = 0.45

    # ---- viewer (unused, headless) ----
    def query_viewer_has_closed(self, v):
        return False


_GYM = None


def acquire_gym():
    global _GYM
    _GYM = Gym()
    return _GYM
