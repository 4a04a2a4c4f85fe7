"""Height-field terrains for the rough-terrain path.

`Terrain` / `HumanoidTerrain` keep the reference's constructor, attributes and random-draw order
(reference utils/terrain.py:38-231): `height_field_raw` / `heightsamples` (int16, (tot_rows, tot_cols)),
`env_origins` (num_rows, num_cols, 3), `vertices` / `triangles` for mesh_type 'trimesh', `env_length`,
`env_width`, `border`, `tot_rows`, `tot_cols`.  Host-side numpy, executed once at start-up; per-step use of the
height field (sampling around the robots, the curriculum) is CUDA (csrc/hg_terrain.cu).

The sub-terrain primitives (`SubTerrain`, `random_uniform_terrain`, `pyramid_sloped_terrain`,
`pyramid_stairs_terrain`, `discrete_obstacles_terrain`, `stepping_stones_terrain`,
`convert_heightfield_to_trimesh`) belong to Isaac Gym Preview 4's `isaacgym.terrain_utils`, a third-party
dependency that is neither vendored in the reference nor installable here; they are restated below from its
published behaviour.  No reference test pins them -> PARITY UNPINNED at that seam; the
`Terrain` / `HumanoidTerrain` logic above it is pinned by tests/golden/terrain.npz.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# isaacgym.terrain_utils, restated
# ------------------------------------------------------------------------------------------------
class SubTerrain:
    def __init__(self, terrain_name="terrain", width=256, length=256, vertical_scale=1.0, horizontal_scale=1.0):
        self.terrain_name = terrain_name
        self.vertical_scale, self.horizontal_scale = vertical_scale, horizontal_scale
        self.width, self.length = width, length
        self.height_field_raw = np.zeros((width, length), dtype=np.int16)


def _bilinear_resample(coarse, n_rows, n_cols):
    """Linear interpolation of a regular grid onto n_rows x n_cols points spanning the same extent
    (what scipy's interp2d(kind='linear') gives for regularly spaced samples)."""
    r = np.linspace(0.0, coarse.shape[0] - 1.0, n_rows)
    c = np.linspace(0.0, coarse.shape[1] - 1.0, n_cols)
    r0 = np.clip(np.floor(r).astype(int), 0, coarse.shape[0] - 2)
    c0 = np.clip(np.floor(c).astype(int), 0, coarse.shape[1] - 2)
    fr, fc = (r - r0)[:, None], (c - c0)[None, :]
    a, b = coarse[np.ix_(r0, c0)], coarse[np.ix_(r0, c0 + 1)]
    d, e = coarse[np.ix_(r0 + 1, c0)], coarse[np.ix_(r0 + 1, c0 + 1)]
    return (a * (1 - fc) + b * fc) * (1 - fr) + (d * (1 - fc) + e * fc) * fr


def random_uniform_terrain(terrain, min_height, max_height, step=1, downsampled_scale=None):
    """Heights drawn uniformly from {min, min+step, ..., max} on a coarse grid, linearly up-sampled, ADDED."""
    if downsampled_scale is None:
        downsampled_scale = terrain.horizontal_scale
    lo, hi = int(min_height / terrain.vertical_scale), int(max_height / terrain.vertical_scale)
    q = int(step / terrain.vertical_scale)
    levels = np.arange(lo, hi + q, q)
    coarse = np.random.choice(levels, (int(terrain.width * terrain.horizontal_scale / downsampled_scale),
                                       int(terrain.length * terrain.horizontal_scale / downsampled_scale)))
    fine = np.rint(_bilinear_resample(coarse.astype(np.float64), terrain.width, terrain.length))
    terrain.height_field_raw += fine.astype(np.int16)
    return terrain


def pyramid_sloped_terrain(terrain, slope=1, platform_size=1.0):
    """Pyramid of the given slope (negative: a bowl) with a flat top of platform_size metres, ADDED then clipped
    at the platform height."""
    w, l = terrain.width, terrain.length
    cx, cy = int(w / 2), int(l / 2)
    fx = ((cx - np.abs(cx - np.arange(w))) / cx).reshape(w, 1)
    fy = ((cy - np.abs(cy - np.arange(l))) / cy).reshape(1, l)
    peak = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * (w / 2))
    terrain.height_field_raw += (peak * fx * fy).astype(terrain.height_field_raw.dtype)
    half = int(platform_size / terrain.horizontal_scale / 2)
    edge = terrain.height_field_raw[w // 2 - half, l // 2 - half]
    terrain.height_field_raw = np.clip(terrain.height_field_raw, min(edge, 0), max(edge, 0))
    return terrain


def pyramid_stairs_terrain(terrain, step_width, step_height, platform_size=1.0):
    """Concentric square steps rising (or descending) towards a central platform; SETS heights."""
    sw = int(step_width / terrain.horizontal_scale)
    sh = int(step_height / terrain.vertical_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    x0, x1, y0, y1, h = 0, terrain.width, 0, terrain.length, 0
    while (x1 - x0) > plat and (y1 - y0) > plat:
        x0, x1, y0, y1, h = x0 + sw, x1 - sw, y0 + sw, y1 - sw, h + sh
        terrain.height_field_raw[x0:x1, y0:y1] = h
    return terrain


def discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size=1.0):
    """num_rects axis-aligned boxes of random size / position / height (+-max, +-max/2); flat central platform."""
    mh = int(max_height / terrain.vertical_scale)
    lo, hi = int(min_size / terrain.horizontal_scale), int(max_size / terrain.horizontal_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    rows, cols = terrain.height_field_raw.shape
    heights = [-mh, -mh // 2, mh // 2, mh]
    sizes = range(lo, hi, 4)
    for _ in range(num_rects):
        w = np.random.choice(sizes)
        l = np.random.choice(sizes)
        i0 = np.random.choice(range(0, rows - w, 4))
        j0 = np.random.choice(range(0, cols - l, 4))
        terrain.height_field_raw[i0:i0 + w, j0:j0 + l] = np.random.choice(heights)
    x0, x1 = (terrain.width - plat) // 2, (terrain.width + plat) // 2
    y0, y1 = (terrain.length - plat) // 2, (terrain.length + plat) // 2
    terrain.height_field_raw[x0:x1, y0:y1] = 0
    return terrain


def stepping_stones_terrain(terrain, stone_size, stone_distance, max_height, platform_size=1.0, depth=-10):
    """Square stones of random height separated by `depth`-deep gaps, laid out in staggered strips."""
    size = int(stone_size / terrain.horizontal_scale)
    gap = int(stone_distance / terrain.horizontal_scale)
    mh = int(max_height / terrain.vertical_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    levels = np.arange(-mh - 1, mh, step=1)
    hf = terrain.height_field_raw
    hf[:, :] = int(depth / terrain.vertical_scale)
    if terrain.length >= terrain.width:
        y = 0
        while y < terrain.length:
            y1 = min(terrain.length, y + size)
            x = np.random.randint(0, size)
            hf[0:max(0, x - gap), y:y1] = np.random.choice(levels)
            while x < terrain.width:
                hf[x:min(terrain.width, x + size), y:y1] = np.random.choice(levels)
                x += size + gap
            y += size + gap
    else:
        x = 0
        while x < terrain.width:
            x1 = min(terrain.width, x + size)
            y = np.random.randint(0, size)
            hf[x:x1, 0:max(0, y - gap)] = np.random.choice(levels)
            while y < terrain.length:
                hf[x:x1, y:min(terrain.length, y + size)] = np.random.choice(levels)
                y += size + gap
            x += size + gap
    x0, x1 = (terrain.width - plat) // 2, (terrain.width + plat) // 2
    y0, y1 = (terrain.length - plat) // 2, (terrain.length + plat) // 2
    hf[x0:x1, y0:y1] = 0
    return terrain


def convert_heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    """Two triangles per grid cell; where the height step between neighbours exceeds slope_threshold the lower
    vertex is moved under the upper one so that the mesh shows a vertical face instead of a steep ramp."""
    hf = height_field_raw
    nr, nc = hf.shape
    yy, xx = np.meshgrid(np.linspace(0, (nc - 1) * horizontal_scale, nc), np.linspace(0, (nr - 1) * horizontal_scale, nr))
    if slope_threshold is not None:
        thr = slope_threshold * horizontal_scale / vertical_scale
        mx, my, mc = np.zeros((nr, nc)), np.zeros((nr, nc)), np.zeros((nr, nc))
        mx[:-1, :] += hf[1:, :] - hf[:-1, :] > thr
        mx[1:, :] -= hf[:-1, :] - hf[1:, :] > thr
        my[:, :-1] += hf[:, 1:] - hf[:, :-1] > thr
        my[:, 1:] -= hf[:, :-1] - hf[:, 1:] > thr
        mc[:-1, :-1] += hf[1:, 1:] - hf[:-1, :-1] > thr
        mc[1:, 1:] -= hf[:-1, :-1] - hf[1:, 1:] > thr
        xx += (mx + mc * (mx == 0)) * horizontal_scale
        yy += (my + mc * (my == 0)) * horizontal_scale
    vertices = np.zeros((nr * nc, 3), dtype=np.float32)
    vertices[:, 0], vertices[:, 1], vertices[:, 2] = xx.flatten(), yy.flatten(), hf.flatten() * vertical_scale
    # cell (i, j): corners v0 = i*nc + j, v1 = v0 + 1, v2 = v0 + nc, v3 = v2 + 1 -> triangles (v0, v3, v1), (v0, v2, v3)
    v0 = (np.arange(nr - 1)[:, None] * nc + np.arange(nc - 1)[None, :]).reshape(-1)
    tri = np.empty((v0.size, 2, 3), dtype=np.uint32)
    tri[:, 0, 0], tri[:, 0, 1], tri[:, 0, 2] = v0, v0 + nc + 1, v0 + 1
    tri[:, 1, 0], tri[:, 1, 1], tri[:, 1, 2] = v0, v0 + nc, v0 + nc + 1
    return vertices, tri.reshape(-1, 3)


# ------------------------------------------------------------------------------------------------
# the reference's own additions (utils/terrain.py:152-176)
# ------------------------------------------------------------------------------------------------
def gap_terrain(terrain, gap_size, platform_size=1.0):
    g = int(gap_size / terrain.horizontal_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    cx, cy = terrain.length // 2, terrain.width // 2
    ix, iy = (terrain.length - plat) // 2, (terrain.width - plat) // 2
    ox, oy = ix + g, iy + g
    terrain.height_field_raw[cx - ox:cx + ox, cy - oy:cy + oy] = -1000
    terrain.height_field_raw[cx - ix:cx + ix, cy - iy:cy + iy] = 0


def pit_terrain(terrain, depth, platform_size=1.0):
    d = int(depth / terrain.vertical_scale)
    half = int(platform_size / terrain.horizontal_scale / 2)
    cx, cy = terrain.length // 2, terrain.width // 2
    terrain.height_field_raw[cx - half:cx + half, cy - half:cy + half] = -d


_PRIMITIVES = dict(random_uniform_terrain=random_uniform_terrain, pyramid_sloped_terrain=pyramid_sloped_terrain,
                   pyramid_stairs_terrain=pyramid_stairs_terrain, discrete_obstacles_terrain=discrete_obstacles_terrain,
                   stepping_stones_terrain=stepping_stones_terrain, gap_terrain=gap_terrain, pit_terrain=pit_terrain)


class Terrain:
    """Grid of num_rows (difficulty levels) x num_cols (terrain types) sub-terrains inside a flat border
    (reference utils/terrain.py:38-150)."""

    def __init__(self, cfg, num_robots):
        self.cfg, self.num_robots, self.type = cfg, num_robots, cfg.mesh_type
        if self.type in ("none", "plane"):
            return
        self.env_length, self.env_width = cfg.terrain_length, cfg.terrain_width
        self.proportions = [np.sum(cfg.terrain_proportions[:i + 1]) for i in range(len(cfg.terrain_proportions))]
        cfg.num_sub_terrains = cfg.num_rows * cfg.num_cols
        self.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        self.width_per_env_pixels = int(self.env_width / cfg.horizontal_scale)
        self.length_per_env_pixels = int(self.env_length / cfg.horizontal_scale)
        self.border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_cols = int(cfg.num_cols * self.width_per_env_pixels) + 2 * self.border
        self.tot_rows = int(cfg.num_rows * self.length_per_env_pixels) + 2 * self.border
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        if cfg.curriculum:
            self.curiculum()
        elif cfg.selected:
            self.selected_terrain()
        else:
            self.randomized_terrain()
        self.heightsamples = self.height_field_raw
        if self.type == "trimesh":
            self.vertices, self.triangles = convert_heightfield_to_trimesh(
                self.height_field_raw, cfg.horizontal_scale, cfg.vertical_scale, cfg.slope_treshold)

    # -- how the grid is populated ---------------------------------------------------------------
    def _random_difficulty(self):                      # :78
        return np.random.choice([0.5, 0.75, 0.9])

    def randomized_terrain(self):                      # :71-80: row-major over the cells, two draws per cell
        for k in range(self.cfg.num_sub_terrains):
            i, j = np.unravel_index(k, (self.cfg.num_rows, self.cfg.num_cols))
            choice = np.random.uniform(0, 1)
            self.add_terrain_to_map(self.make_terrain(choice, self._random_difficulty()), i, j)

    def curiculum(self):                               # :82-90: column by column, difficulty grows with the row
        for j in range(self.cfg.num_cols):
            for i in range(self.cfg.num_rows):
                self.add_terrain_to_map(self.make_terrain(j / self.cfg.num_cols + 0.001, i / self.cfg.num_rows), i, j)

    def selected_terrain(self):                        # :92-105 (one primitive, named in terrain_kwargs['type'], everywhere)
        kwargs = dict(self.cfg.terrain_kwargs)
        fn = _PRIMITIVES[kwargs.pop("type").split(".")[-1]]
        kwargs = kwargs.get("terrain_kwargs", kwargs)
        for k in range(self.cfg.num_sub_terrains):
            i, j = np.unravel_index(k, (self.cfg.num_rows, self.cfg.num_cols))
            t = self._blank()
            fn(t, **kwargs)
            self.add_terrain_to_map(t, i, j)

    def _blank(self):
        return SubTerrain("terrain", width=self.width_per_env_pixels, length=self.width_per_env_pixels,
                          vertical_scale=self.cfg.vertical_scale, horizontal_scale=self.cfg.horizontal_scale)

    def make_terrain(self, choice, difficulty):        # :107-146
        t, p = self._blank(), self.proportions
        slope, step_h = difficulty * 0.4, 0.05 + 0.18 * difficulty
        if choice < p[0]:
            pyramid_sloped_terrain(t, slope=-slope if choice < p[0] / 2 else slope, platform_size=3.0)
        elif choice < p[1]:
            pyramid_sloped_terrain(t, slope=slope, platform_size=3.0)
            random_uniform_terrain(t, min_height=-0.05, max_height=0.05, step=0.005, downsampled_scale=0.2)
        elif choice < p[3]:
            pyramid_stairs_terrain(t, step_width=0.31, step_height=-step_h if choice < p[2] else step_h, platform_size=3.0)
        elif choice < p[4]:
            discrete_obstacles_terrain(t, 0.05 + difficulty * 0.2, 1.0, 2.0, 20, platform_size=3.0)
        elif choice < p[5]:
            stepping_stones_terrain(t, stone_size=1.5 * (1.05 - difficulty), stone_distance=0.05 if difficulty == 0 else 0.1,
                                    max_height=0.0, platform_size=4.0)
        elif choice < p[6]:
            gap_terrain(t, gap_size=1.0 * difficulty, platform_size=3.0)
        else:
            pit_terrain(t, depth=1.0 * difficulty, platform_size=4.0)
        return t

    def add_terrain_to_map(self, terrain, row, col):   # :148-166
        L, W, b = self.length_per_env_pixels, self.width_per_env_pixels, self.border
        self.height_field_raw[b + row * L:b + (row + 1) * L, b + col * W:b + (col + 1) * W] = terrain.height_field_raw
        hs = terrain.horizontal_scale
        x0, x1 = int((self.env_length / 2.0 - 1) / hs), int((self.env_length / 2.0 + 1) / hs)
        y0, y1 = int((self.env_width / 2.0 - 1) / hs), int((self.env_width / 2.0 + 1) / hs)
        top = np.max(terrain.height_field_raw[x0:x1, y0:y1]) * terrain.vertical_scale      # spawn on the highest point near the centre
        self.env_origins[row, col] = [(row + 0.5) * self.env_length, (col + 0.5) * self.env_width, top]


class HumanoidTerrain(Terrain):
    """The XBot-L terrain mix: plane, obstacles, uniform noise, slope up / down, stairs up / down
    (reference utils/terrain.py:189-231)."""

    def _random_difficulty(self):                      # :198
        return np.random.uniform(0, 1)

    def make_terrain(self, choice, difficulty):        # :202-231
        t, p = self._blank(), self.proportions
        obstacle_h, noise_h, slope = difficulty * 0.04, difficulty * 0.07, difficulty * 0.15
        if choice < p[0]:
            pass
        elif choice < p[1]:
            discrete_obstacles_terrain(t, obstacle_h, 1.0, 2.0, 20, platform_size=3.0)
        elif choice < p[2]:
            random_uniform_terrain(t, min_height=-noise_h, max_height=noise_h, step=0.005, downsampled_scale=0.2)
        elif choice < p[3]:
            pyramid_sloped_terrain(t, slope=slope, platform_size=0.1)
        elif choice < p[4]:
            pyramid_sloped_terrain(t, slope=-slope, platform_size=0.1)
        elif choice < p[5]:
            pyramid_stairs_terrain(t, step_width=0.4, step_height=obstacle_h, platform_size=1.0)
        elif choice < p[6]:
            pyramid_stairs_terrain(t, step_width=0.4, step_height=-obstacle_h, platform_size=1.0)
        return t
