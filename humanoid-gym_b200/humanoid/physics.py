"""The physics seam below the hot path ("Physics boundary").

The reference talks to Isaac Gym through its tensor API: four state tensors
(`actor_root_state (N,13)`, `dof_state (N*12,2)`, `net_contact_force (N*13,3)`,
`rigid_body_state (N*13,13)`), `simulate`, `refresh_*`, `set_*` (reference
legged_robot.py:96-101,124-126,371-373,395-397,438-457; humanoid_env.py:97-98).
`PhysicsBackend` keeps exactly that surface so `LeggedRobot` reads like the reference:

* `IsaacGymPhysics`   -- the adapter over the real simulator (humanoid/isaacgym_physics.py), used when `isaacgym` is
  importable (not in this image: the tests run it over the functional fake in tests/golden/fake_isaacgym).
* `SyntheticPhysics`  -- seeded synthetic tensor source: a ring of
  pre-generated frames, either resident in HBM or in pinned host memory (`host_resident=True`,
  the end-to-end bench arm: every refresh is then a host->device copy inside the step).

Nothing here is on the accelerated path; it is the producer of the tensors the fused env kernel
consumes.
"""
import torch

from humanoid.envs.custom import xbot_l_model as robot
from humanoid.synthetic_frames import generate_ring


class PhysicsBackend:
    """Isaac-Gym-tensor-API shaped interface."""
    num_envs: int
    device: torch.device
    root_states: torch.Tensor      # (N, 13)
    dof_state: torch.Tensor        # (N*num_dof, 2)
    contact_forces: torch.Tensor   # (N*num_bodies, 3)
    rigid_state: torch.Tensor      # (N*num_bodies, 13)

    num_dof = robot.NUM_DOF
    num_bodies = robot.NUM_BODIES
    dof_names = robot.DOF_NAMES
    body_names = robot.BODY_NAMES

    def dof_properties(self):
        return dict(lower=list(robot.DOF_LOWER), upper=list(robot.DOF_UPPER),
                    velocity=list(robot.DOF_VELOCITY), effort=list(robot.DOF_EFFORT))

    def add_terrain(self, terrain, mesh_type):
        """gym.add_heightfield / gym.add_triangle_mesh (reference legged_robot.py:553-586): `terrain` is the
        utils.terrain.Terrain whose heightsamples / vertices / triangles a simulator collides the robots with.
        Sources that do not simulate contact ignore it."""
        pass

    # -- stepping ---------------------------------------------------------------------------
    def set_dof_actuation_force_tensor(self, torques):
        pass

    def simulate(self):
        raise NotImplementedError

    def fetch_results(self):
        pass

    def refresh_dof_state_tensor(self):
        pass

    def refresh_actor_root_state_tensor(self):
        pass

    def refresh_net_contact_force_tensor(self):
        pass

    def refresh_rigid_body_state_tensor(self):
        pass

    # -- state injection (resets / pushes) ----------------------------------------------------
    def set_dof_state_tensor_indexed(self, dof_state, env_ids_int32, count):
        pass

    def set_actor_root_state_tensor(self, root_states):
        pass

    def set_actor_root_state_tensor_indexed(self, root_states, env_ids_int32, count):
        pass

    def fused_decimation(self, env):
        """Optional fast path: run the whole decimation loop of one env step (PD torques, set forces, simulate, refresh
        dof) AND the root / contact / rigid refreshes in one go.  Returns True when done; the default (a real
        simulator) returns False and LeggedRobot.step runs the reference's loop."""
        return False

    def apply_env_writes(self, reset_ids, scratch, pushed):
        """Called once per env step after the fused kernel: `reset_ids[:scratch[3]]` are the envs
        whose dof/root rows were rewritten, `pushed` tells whether root velocities were overwritten
        for all envs.  Backends that own a real simulator forward this to set_*_indexed."""
        pass


class SyntheticPhysics(PhysicsBackend):
    """Seeded synthetic stand-in for gym.simulate() + refresh_* (SURVEY.md section 8d).

    K frames are generated once (outside any timed region).  `simulate()` only advances a counter;
    `refresh_*` copy the current frame into the live state tensors, which is what PhysX does to its
    GPU buffers from the env's point of view."""

    def __init__(self, num_envs, device, cmd_ranges, env_origins=None, decimation=10, seed=5, ring=6,
                 host_resident=False, p_base_contact=0.002):
        self.num_envs, self.device = num_envs, torch.device(device)
        self.decimation, self.ring = decimation, ring
        self.host_resident = host_resident
        N, nb, nd = num_envs, self.num_bodies, self.num_dof
        dev = self.device
        K = ring
        fr = generate_ring(N, dev, cmd_ranges, env_origins, robot.DOF_LOWER, robot.DOF_UPPER, nb, decimation=decimation,
                           seed=seed, ring=ring, p_base_contact=p_base_contact)
        root, dof, contact, rigid = fr["root"], fr["dof"], fr["contact"], fr["rigid"]

        def place(t):
            return t.cpu().pin_memory() if host_resident else t

        self._ring_root, self._ring_dof = place(root), place(dof.view(K * decimation, N * nd, 2))
        self._ring_contact, self._ring_rigid = place(contact.view(K, N * nb, 3)), place(rigid.view(K, N * nb, 13))
        # live tensors handed out through acquire_*
        self.root_states = torch.zeros(N, 13, device=dev)
        self.root_states[:, 6] = 1.0
        self.root_states[:, 2] = 0.95
        self.dof_state = torch.zeros(N * nd, 2, device=dev)
        self.contact_forces = torch.zeros(N * nb, 3, device=dev)
        self.rigid_state = torch.zeros(N * nb, 13, device=dev)
        self.substep = 0
        self.h2d_bytes = 0
        import os
        self.fused = os.environ.get("HG_FUSED_DECIMATION", "1") != "0"      # 0: the reference's decimation loop, launch by launch
        # host-resident frames are staged through a double buffer in HBM by a copy stream: the frames of env step
        # s+1 cross PCIe while step s computes (the synthetic source is open-loop, so the next frames are known)
        self._rollout_left = None            # env steps left in the rollout announced by begin_rollout()
        if host_resident:
            self._stage = [dict(dof=torch.empty(decimation, N * nd, 2, device=dev), root=torch.empty(N, 13, device=dev),
                                contact=torch.empty(N * nb, 3, device=dev), rigid=torch.empty(N * nb, 13, device=dev))
                           for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(dev)
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._free = [torch.cuda.Event(), torch.cuda.Event()]
            self._staged_step = [-1, -1]     # which env step each slot holds (host mirror)
            self._ready_pending = [False, False]
            self._free_recorded = [False, False]
            self._cur_step = -1

    # ---- host-resident staging ---------------------------------------------------------------------------
    def _enqueue_stage(self, step, stream):
        """Copy the frames of env step `step` (ring indices follow from it) into slot step % 2 on `stream`."""
        slot = step % 2
        st = self._stage[slot]
        k = step % self.ring
        d0 = (step * self.decimation) % (self.ring * self.decimation)
        with torch.cuda.stream(stream):
            st["dof"].copy_(self._ring_dof[d0:d0 + self.decimation], non_blocking=True)
            st["root"].copy_(self._ring_root[k], non_blocking=True)
            st["contact"].copy_(self._ring_contact[k], non_blocking=True)
            st["rigid"].copy_(self._ring_rigid[k], non_blocking=True)
        self._staged_step[slot] = step

    def begin_rollout(self, steps):
        """The next `steps` env steps form one collection phase (possibly being captured into a CUDA graph): its first
        step loads its frames in line, its last step does not prefetch, and no event recorded before this call is
        waited on afterwards -- so the phase is self-contained and can be replayed."""
        self._rollout_left = int(steps)
        if self.host_resident:
            self._ready_pending = [False, False]
            self._free_recorded = [False, False]
            self._cur_step = -1

    def _begin_step(self, step):
        """Called at the first substep of env step `step`: make its frames available, start fetching the next."""
        cur = torch.cuda.current_stream(self.device)
        slot = step % 2
        if self._cur_step >= 0:
            self._free[(step - 1) % 2].record(cur)           # every read of the previous step's slot is enqueued by now
            self._free_recorded[(step - 1) % 2] = True
        if not (self._staged_step[slot] == step and self._ready_pending[slot]):
            self._enqueue_stage(step, cur)                   # not prefetched (first step of a rollout): load in line
        else:
            cur.wait_event(self._ready[slot])
        self._ready_pending[slot] = False
        self._cur_step = step
        nxt = step + 1
        if self._rollout_left is not None:
            self._rollout_left -= 1
            if self._rollout_left <= 0:
                self._rollout_left = None
                return                                       # the next rollout loads its first step itself
        cp = self._copy_stream
        cp.wait_event(self._free[nxt % 2]) if self._free_recorded[nxt % 2] else cp.wait_stream(cur)
        self._enqueue_stage(nxt, cp)
        self._ready[nxt % 2].record(cp)
        self._ready_pending[nxt % 2] = True

    # how many bytes one env step moves host->device when host_resident
    def h2d_bytes_per_step(self):
        if not self.host_resident:
            return 0
        per = self._ring_root[0].numel() + self._ring_contact[0].numel() + self._ring_rigid[0].numel()
        return 4 * (per + self.decimation * self._ring_dof[0].numel())

    def _frame(self):
        return ((self.substep - 1) // self.decimation) % self.ring if self.substep > 0 else 0

    def simulate(self):
        self.substep += 1
        if self.host_resident and (self.substep - 1) % self.decimation == 0:
            self._begin_step((self.substep - 1) // self.decimation)

    def _slot(self):
        return self._stage[((self.substep - 1) // self.decimation) % 2]

    def fused_decimation(self, env):
        """One launch (hg_env_synth_decimation) for the `decimation` sub-steps + the three refreshes: the frames are known
        in advance (open loop), so only the arithmetic of the loop remains -- ~23 graph nodes per env step less."""
        if not self.fused or env.cfg.control.decimation != self.decimation:
            return False
        from humanoid import _native as nat
        dec = self.decimation
        step = self.substep // dec
        self.substep += dec
        if self.host_resident:
            self._begin_step(step)
            st = self._stage[step % 2]
            dof, root, contact, rigid = st["dof"], st["root"], st["contact"], st["rigid"]
        else:
            k = step % self.ring
            d0 = (step * dec) % (self.ring * dec)
            dof, root, contact, rigid = self._ring_dof[d0:d0 + dec], self._ring_root[k], self._ring_contact[k], self._ring_rigid[k]
        nat.check(nat.lib.hg_env_synth_decimation(env._B, env._P, dof.data_ptr(), dec, root.data_ptr(), contact.data_ptr(),
                                                  rigid.data_ptr(), self.num_envs, nat.stream_ptr(self.device.index)),
                  "hg_env_synth_decimation")
        return True

    def refresh_dof_state_tensor(self):
        if self.host_resident and self.substep > 0:
            self.dof_state.copy_(self._slot()["dof"][(self.substep - 1) % self.decimation], non_blocking=True)
        else:
            self.dof_state.copy_(self._ring_dof[(self.substep - 1) % (self.ring * self.decimation)], non_blocking=True)

    def refresh_actor_root_state_tensor(self):
        src = self._slot()["root"] if (self.host_resident and self.substep > 0) else self._ring_root[self._frame()]
        self.root_states.copy_(src, non_blocking=True)

    def refresh_net_contact_force_tensor(self):
        src = self._slot()["contact"] if (self.host_resident and self.substep > 0) else self._ring_contact[self._frame()]
        self.contact_forces.copy_(src, non_blocking=True)

    def refresh_rigid_body_state_tensor(self):
        src = self._slot()["rigid"] if (self.host_resident and self.substep > 0) else self._ring_rigid[self._frame()]
        self.rigid_state.copy_(src, non_blocking=True)


class ExternalPhysics(PhysicsBackend):
    """State tensors written by the caller (parity tests: frames come from golden vectors)."""

    def __init__(self, num_envs, device):
        self.num_envs, self.device = num_envs, torch.device(device)
        N, nb, nd = num_envs, self.num_bodies, self.num_dof
        self.root_states = torch.zeros(N, 13, device=device)
        self.root_states[:, 6] = 1.0
        self.dof_state = torch.zeros(N * nd, 2, device=device)
        self.contact_forces = torch.zeros(N * nb, 3, device=device)
        self.rigid_state = torch.zeros(N * nb, 13, device=device)
        self.on_simulate = None

    def simulate(self):
        if self.on_simulate is not None:
            self.on_simulate(self)


def isaacgym_available():
    try:
        import isaacgym  # noqa: F401
        return True
    except Exception:
        return False


def make_physics(kind, num_envs, device, cfg, env_origins, seed=5, rank=0, sim_params=None, physics_engine=None,
                 sim_device_id=0, custom_origins=False):
    """kind: 'auto' | 'synthetic' | 'synthetic_host' | 'external' | 'isaacgym'."""
    if kind == "auto":
        kind = "isaacgym" if isaacgym_available() else "synthetic"
    ranges = {k: getattr(cfg.commands.ranges, k) for k in ("lin_vel_x", "lin_vel_y")}
    if kind in ("synthetic", "synthetic_host"):
        return SyntheticPhysics(num_envs, device, ranges, env_origins, decimation=cfg.control.decimation,
                                seed=seed + rank, host_resident=(kind == "synthetic_host"))
    if kind == "external":
        return ExternalPhysics(num_envs, device)
    if kind == "isaacgym":
        if not isaacgym_available():
            raise RuntimeError(
                "HG_PHYSICS=isaacgym: `import isaacgym` failed (Isaac Gym Preview 4 ships no sm_100 build and is not in this "
                "image) -- select HG_PHYSICS=synthetic, or put a Blackwell-capable build on PYTHONPATH (INTEGRATION.md)")
        from humanoid.isaacgym_physics import IsaacGymPhysics
        return IsaacGymPhysics(num_envs, device, cfg, env_origins, sim_params, physics_engine=physics_engine,
                               sim_device_id=sim_device_id, custom_origins=custom_origins)
    raise ValueError(f"unknown physics backend {kind!r}")
