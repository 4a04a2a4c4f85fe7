"""CPU-side checks of the C-ABI boundary: the shared library loads, exports every symbol declared in
include/hg_b200.h, and the ctypes structs have the library's sizes.  No compute calls (no GPU here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "hg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", text)))


def test_build_and_exports():
    import __graft_entry__ as ge
    lib_path = ge.build()
    assert os.path.exists(lib_path)
    import ctypes
    lib = ctypes.CDLL(lib_path)
    declared = _declared()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in hg_b200.h but not exported"
    from humanoid import _native as nat
    assert set(nat.EXPORTS) == set(declared), set(nat.EXPORTS) ^ set(declared)


def test_struct_sizes_match():
    from humanoid import _native as nat
    import ctypes
    for k, st in enumerate(nat._STRUCTS):
        assert nat.lib.hg_struct_size(k) == ctypes.sizeof(st)
    assert nat.lib.hg_struct_size(99) == -1


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: constructing the env / learner on a CPU device raises."""
    import pytest
    import torch
    from humanoid import _native as nat
    from humanoid.algo import ActorCritic, PPO, RolloutStorage
    with pytest.raises(nat.NativeError):
        RolloutStorage(4, 2, [705], [219], [12], "cpu")
    ac = ActorCritic(47, 73, 12, [8], [8])
    with pytest.raises(nat.NativeError):
        PPO(ac, device="cpu")
    if not torch.cuda.is_available():
        from parity_utils import make_env
        with pytest.raises(Exception):
            make_env(4, device="cpu")


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (tier rule 3)."""
    pkg = os.path.join(ROOT, "humanoid-gym_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)


def test_argument_errors_are_reported_before_any_launch():
    """SURVEY.md 8b error contract: negative HG_E_* codes for NULL pointers / bad sizes / bad arguments, detected on the host
    before any launch (so this runs without a GPU), message through hg_last_error(); nothing throws or aborts."""
    from humanoid import _native as nat
    lib = nat.lib
    E_NULL, E_SIZE, E_ARG = -1, -3, -4
    T = nat.Terrain()
    fake = 0x1000                                        # never dereferenced: every call below fails its host-side checks
    assert lib.hg_terrain_get_heights(None, fake, fake, 187, fake, 4, None) == E_NULL
    assert lib.hg_terrain_get_heights(T, fake, fake, 187, fake, 4, None) == E_NULL           # height_samples is NULL
    assert b"height_samples" in lib.hg_last_error()
    T.height_samples, T.rows, T.cols, T.horizontal_scale = fake, 1, 1, 0.1
    assert lib.hg_terrain_get_heights(T, fake, fake, 187, fake, 4, None) == E_SIZE           # field smaller than 2 x 2
    T.rows, T.cols, T.horizontal_scale = 100, 100, 0.0
    assert lib.hg_terrain_get_heights(T, fake, fake, 187, fake, 4, None) == E_ARG
    T.horizontal_scale = 0.1
    assert lib.hg_terrain_get_heights(T, fake, fake, 0, fake, 4, None) == E_SIZE
    assert lib.hg_terrain_get_heights(T, fake, fake, 187, fake, 0, None) == E_SIZE
    assert lib.hg_terrain_get_heights(T, None, fake, 187, fake, 4, None) == E_NULL
    assert lib.hg_terrain_reset_prepare(T, fake, fake, fake, fake, fake, fake, fake, None, None, 0, 0, None, 4, None) == E_NULL  # origins
    T.terrain_origins, T.num_levels, T.num_types = fake, 4, 5
    assert lib.hg_terrain_reset_prepare(T, fake, fake, fake, fake, fake, fake, fake, None, None, 0, 0, None, 4, None) == E_ARG   # spawn aliases env_origins
    assert b"alias" in lib.hg_last_error()
    assert lib.hg_terrain_reset_prepare(T, None, fake, fake, fake, fake, fake, fake + 64, None, None, 0, 0, None, 4, None) == E_NULL
    assert lib.hg_terrain_priv_frames(fake, 736, 705, fake, fake, 187, 5.0, 18.0, None, fake, fake, 2688, 3, 4, None) == E_ARG   # in == out
    assert lib.hg_terrain_priv_frames(fake, 736, 705, fake, fake, 187, 5.0, 18.0, None, fake, fake + 64, 2000, 3, 4, None) == E_SIZE
    assert lib.hg_terrain_priv_frames(fake, 700, 705, fake, fake, 187, 5.0, 18.0, None, fake, fake + 64, 2688, 3, 4, None) == E_SIZE
    assert lib.hg_randperm(0, 1, 0, fake, None) == E_SIZE and lib.hg_randperm(8, 1, 0, None, None) == E_NULL
    assert lib.hg_clip_adam_step_stats(fake, fake, fake, fake, fake, 1.0, fake, fake, 0.9, 0.999, 1e-8, 1.0, 8, None, fake, 8, None) == E_ARG
    assert lib.hg_gae(None, fake, 0.99, 0.9, fake, 1, 4, None) == E_NULL
    B, P, Z = nat.EnvBuffers(), nat.EnvParams(), nat.EnvNoise()
    assert lib.hg_env_post_physics(B, P, Z, 0x7F, 1, 4, None) == E_NULL
    assert lib.hg_env_pre_physics(B, P, None, None, None, 0, 0, 4, None) == E_NULL
    prev = lib.hg_set_gae_mode(1)
    assert lib.hg_set_gae_mode(prev) == 1
