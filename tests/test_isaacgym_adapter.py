"""Host logic of the real-physics backend (humanoid/isaacgym_physics.py) over the test-only
functional fake `isaacgym` of tests/golden/fake_isaacgym -- Isaac Gym itself ships no sm_100 build and is not in the
image.  CPU only: the adapter is plain gym-API plumbing; the env on top of it is covered by
tests/test_env_gpu.py::test_env_over_isaacgym_adapter."""
import numpy as np
import pytest
import torch

@pytest.fixture()
def fake_gym(fake_isaacgym):
    fake_isaacgym.setenv("HG_FAKE_GYM", "golden")
    yield


def _log_calls(gym, log):
    for name in ("set_dof_actuation_force_tensor", "simulate", "fetch_results", "refresh_dof_state_tensor",
                 "refresh_actor_root_state_tensor", "refresh_net_contact_force_tensor", "refresh_rigid_body_state_tensor",
                 "set_dof_state_tensor_indexed", "set_actor_root_state_tensor", "set_actor_root_state_tensor_indexed",
                 "add_ground", "add_triangle_mesh", "add_heightfield"):
        real = getattr(gym, name)

        def wrapped(*a, _real=real, _name=name):
            log.append((_name, a[1:]))
            return _real(*a)
        setattr(gym, name, wrapped)


def _cfg(mesh_type="plane"):
    from humanoid.envs import XBotLCfg

    class Cfg(XBotLCfg):
        class terrain(XBotLCfg.terrain):
            num_rows, num_cols, border_size = 3, 4, 2
    Cfg.terrain.mesh_type = mesh_type
    return Cfg()


def _sim_params():
    from humanoid.utils.helpers import SimParams
    sp = SimParams()
    sp.dt, sp.use_gpu_pipeline = 0.001, False
    sp.physx.num_position_iterations = 4
    return sp


def test_make_physics_refuses_without_isaacgym():
    from humanoid import physics
    if physics.isaacgym_available():
        pytest.skip("isaacgym importable here")
    with pytest.raises(RuntimeError, match="import isaacgym"):
        physics.make_physics("isaacgym", 4, "cpu", _cfg(), torch.zeros(4, 3))


@pytest.mark.parametrize("mesh_type", ["plane", "trimesh", "heightfield"])
def test_adapter_builds_sim_like_the_reference(fake_gym, mesh_type):
    from humanoid import physics
    from humanoid.utils.terrain import HumanoidTerrain
    assert physics.isaacgym_available()
    N = 12
    cfg = _cfg(mesh_type)
    cfg.domain_rand.randomize_base_mass = True
    origins = torch.arange(N * 3, dtype=torch.float32).view(N, 3)
    torch.manual_seed(0)
    np.random.seed(0)
    ph = physics.make_physics("isaacgym", N, "cpu", cfg, origins, sim_params=_sim_params(), physics_engine=1, sim_device_id=0)
    log = []
    _log_calls(ph.gym, log)
    assert ph._gym_params.dt == float(np.float32(0.001)) and ph._gym_params.physx.num_position_iterations == 4
    assert (ph.num_dof, ph.num_bodies) == (12, 13) and ph.dof_names[0].startswith("left") and ph.body_names[0] == "base_link"
    props = ph.dof_properties()
    assert len(props["lower"]) == 12 and all(lo < hi for lo, hi in zip(props["lower"], props["upper"]))
    if mesh_type != "plane":
        ph.add_terrain(HumanoidTerrain(cfg.terrain, N), mesh_type)
    fr, mass = ph.create_actors()
    ph.prepare()
    kinds = [c[0] for c in log]
    assert kinds.count({"plane": "add_ground", "trimesh": "add_triangle_mesh", "heightfield": "add_heightfield"}[mesh_type]) == 1
    assert len(ph.envs) == N and ph.body_index(cfg.asset.foot_name and "left_ankle_roll_link") == 6
    lo, hi = cfg.domain_rand.friction_range
    assert fr.shape == (N, 1) and float(fr.min()) >= lo and float(fr.max()) <= hi and ph.friction_coeffs.shape == (N, 1, 1)
    a, b = cfg.domain_rand.added_mass_range
    assert mass.shape == (N, 1) and float((mass - 5.0).min()) >= a and float((mass - 5.0).max()) <= b and float(mass.std()) > 0
    # actors start within 1 m (xy) of their origin
    placed = torch.tensor(ph.sim.origins)
    assert float((placed[:, :2] - origins[:, :2]).abs().max()) <= 1.0 and torch.equal(placed[:, 2], origins[:, 2])
    assert ph.root_states.shape == (N, 13) and ph.dof_state.shape == (N * 12, 2)
    assert ph.contact_forces.shape == (N * 13, 3) and ph.rigid_state.shape == (N * 13, 13)

    # one decimation sub-step + the post-physics refreshes, then the state writes of a reset / push
    log.clear()
    torques = torch.zeros(N, 12)
    ph.set_dof_actuation_force_tensor(torques)
    ph.simulate()
    ph.refresh_dof_state_tensor()
    ph.refresh_actor_root_state_tensor()
    ph.refresh_net_contact_force_tensor()
    ph.refresh_rigid_body_state_tensor()
    assert [c[0] for c in log] == ["set_dof_actuation_force_tensor", "simulate", "fetch_results", "refresh_dof_state_tensor",
                                   "refresh_actor_root_state_tensor", "refresh_net_contact_force_tensor",
                                   "refresh_rigid_body_state_tensor"]
    assert log[0][1][0] is torques and float(ph.root_states[:, 3:7].norm(dim=1).min()) > 0.99    # the fake wrote a state
    log.clear()
    reset_ids = torch.tensor([7, 2, 9, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=torch.int32)
    scratch = torch.zeros(32, dtype=torch.int32)
    scratch[3] = 3
    ph.apply_env_writes(reset_ids, scratch, pushed=True)
    assert [c[0] for c in log] == ["set_actor_root_state_tensor", "set_dof_state_tensor_indexed", "set_actor_root_state_tensor_indexed"]
    assert log[1][1][0] is ph.dof_state and log[1][1][1].tolist() == [7, 2, 9] and log[1][1][2] == 3
    assert log[2][1][0] is ph.root_states and log[2][1][1].dtype == torch.int32
    log.clear()
    scratch[3] = 0
    ph.apply_env_writes(reset_ids, scratch, pushed=False)
    assert log == []
