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
