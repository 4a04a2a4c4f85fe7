"""TEST-ONLY stand-in for NVIDIA Isaac Gym Preview 4 (not installable here).

Used only by tests/golden/make_golden.py to import and execute the UNMODIFIED
reference (/root/reference) on CPU so that golden vectors can be captured.
It is not part of the product and is never imported by it.

`simulate()` is a seeded synthetic tensor writer (no physics).  The math
helpers in torch_utils follow Isaac Gym's public definitions (xyzw quats);
no reference test pins them -> "parity unpinned" at this boundary.
"""
from . import gymapi, gymtorch, gymutil, torch_utils, terrain_utils  # noqa: F401
