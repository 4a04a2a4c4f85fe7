"""Host logic of the pinned-host physics source (CPU, CUDA stream / event API faked).

`SyntheticPhysics(host_resident=True)` stages the frames of env step s+1 through a double buffer on a copy stream
while step s runs.  Two properties are checked without a GPU:
  * the live tensors always hold the same frames as the HBM-resident variant (copies run synchronously here);
  * capture safety: after `begin_rollout()` no stream waits on an event whose last record predates it, the last
    step of the announced rollout starts no prefetch, and every prefetch is joined (waited on) before the rollout
    ends -- the conditions under which a rollout can be stream-captured once and replayed."""
import contextlib

import pytest
import torch

from humanoid import physics as phys


class _Clock:
    t = 0

    @classmethod
    def tick(cls):
        cls.t += 1
        return cls.t


class FakeEvent:
    def __init__(self, *a, **k):
        self.recorded_at, self.stream = None, None

    def record(self, stream=None):
        self.recorded_at, self.stream = _Clock.tick(), stream
        stream.log.append(("record", self))


class FakeStream:
    def __init__(self, *a, name="copy", **k):
        self.name, self.log = name, []

    def wait_event(self, ev):
        assert ev.recorded_at is not None, "wait on an event that was never recorded"
        self.log.append(("wait_event", ev, ev.recorded_at, _Clock.tick()))

    def wait_stream(self, other):
        self.log.append(("wait_stream", other, _Clock.tick()))


@pytest.fixture
def fake_cuda(monkeypatch):
    main = FakeStream(name="main")
    active = [main]

    @contextlib.contextmanager
    def stream_ctx(s):
        active.append(s)
        try:
            yield
        finally:
            active.pop()

    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: active[-1])
    monkeypatch.setattr(torch.cuda, "stream", stream_ctx)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    return main


def _make(host, ring=6, dec=10, N=8):
    ranges = {"lin_vel_x": [-0.3, 0.6], "lin_vel_y": [-0.3, 0.3]}
    return phys.SyntheticPhysics(N, "cpu", ranges, decimation=dec, seed=11, ring=ring, host_resident=host)


def _env_step(p):
    """The refresh sequence of one LeggedRobot.step(): dec x (simulate, dof refresh), then root / contact / rigid."""
    frames = []
    for _ in range(p.decimation):
        p.simulate()
        p.refresh_dof_state_tensor()
        frames.append(p.dof_state.clone())
    p.refresh_actor_root_state_tensor()
    p.refresh_net_contact_force_tensor()
    p.refresh_rigid_body_state_tensor()
    return frames + [p.root_states.clone(), p.contact_forces.clone(), p.rigid_state.clone()]


def test_host_resident_source_serves_the_same_frames(fake_cuda):
    a, b = _make(False), _make(True)
    for step in range(15):                          # wraps the ring of 6 twice
        if step in (2, 8):
            b.begin_rollout(6)
        for x, y in zip(_env_step(a), _env_step(b)):
            assert torch.equal(x, y), step
    assert b.h2d_bytes_per_step() == 4 * (8 * 13 + 8 * 13 * 3 + 8 * 13 * 13 + 10 * 8 * 12 * 2)
    assert a.h2d_bytes_per_step() == 0


def test_rollout_is_self_contained_for_stream_capture(fake_cuda):
    main = fake_cuda
    p = _make(True)
    for _ in range(3):                               # eager steps before the rollout leave a prefetch in flight
        _env_step(p)
    copy = p._copy_stream
    T = 6
    t_begin = _Clock.tick()
    n_main, n_copy = len(main.log), len(copy.log)
    p.begin_rollout(T)
    for _ in range(T):
        _env_step(p)
    events = main.log[n_main:] + copy.log[n_copy:]
    waits = [e for e in events if e[0] == "wait_event"]
    assert waits, "the rollout never joined a prefetch"
    for _, ev, recorded_at, _t in waits:             # nothing recorded before begin_rollout() is waited on
        assert recorded_at > t_begin
    # T - 1 prefetches (none on the last step), each recorded on the copy stream and later waited on by the main stream
    ready_records = [e[1] for e in copy.log[n_copy:] if e[0] == "record"]
    assert len(ready_records) == T - 1
    joined = {id(e[1]) for e in main.log[n_main:] if e[0] == "wait_event"}
    assert all(id(ev) in joined for ev in ready_records)
    # the copy stream never runs ahead of the consumer of the slot it overwrites: it waits for a "free" record of
    # the main stream (or for the main stream itself) before every prefetch
    copy_waits = [e for e in copy.log[n_copy:] if e[0] in ("wait_event", "wait_stream")]
    assert len(copy_waits) == T - 1
    assert p._rollout_left is None
