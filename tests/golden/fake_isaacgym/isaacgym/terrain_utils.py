"""TEST-ONLY stand-in for isaacgym.terrain_utils: the product's restatement of the Isaac Gym primitives
(humanoid-gym_b200/humanoid/utils/terrain.py, loaded by file path -- the product package is not imported).  The
reference's Terrain / HumanoidTerrain classes (utils/terrain.py) run UNMODIFIED on top of it, which is what
tests/golden/terrain.npz pins; the primitives themselves are third-party and parity-unpinned."""
import importlib.util
import os

_REL = os.path.join("humanoid-gym_b200", "humanoid", "utils", "terrain.py")
_HERE5 = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
# the repo root: HG_REPO_ROOT, else five levels up (in-tree), else /root/repo (this file copied elsewhere, e.g. to /tmp)
_REPO = next((r for r in (os.environ.get("HG_REPO_ROOT"), _HERE5, "/root/repo") if r and os.path.exists(os.path.join(r, _REL))), _HERE5)
_spec = importlib.util.spec_from_file_location("_hg_terrain_primitives", os.path.join(_REPO, _REL))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)

SubTerrain = _m.SubTerrain
random_uniform_terrain = _m.random_uniform_terrain
pyramid_sloped_terrain = _m.pyramid_sloped_terrain
pyramid_stairs_terrain = _m.pyramid_stairs_terrain
discrete_obstacles_terrain = _m.discrete_obstacles_terrain
stepping_stones_terrain = _m.stepping_stones_terrain
convert_heightfield_to_trimesh = _m.convert_heightfield_to_trimesh
