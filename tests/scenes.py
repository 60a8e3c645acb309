"""Shared synthetic scenes for tests / smoke / bench.  Pure numpy."""
import numpy as np

from claymore_b200 import samplers

J_FLUID, FIXED_COROTATED, SAND, NACC = 0, 1, 2, 3


def jelly_cube(domain_bits=7, lo=51, hi=77):
    """Config 1: 128^3 grid, lattice cube [51,77)^3 cells x 8 = 140 608 particles, v0 = (0,-1,0)."""
    dx = 1.0 / (1 << domain_bits)
    return dict(domain_bits=domain_bits, models=[dict(material=FIXED_COROTATED, pos=samplers.uniform_box(dx, (lo,) * 3, (hi,) * 3), v0=(0.0, -1.0, 0.0))])


def small_cube(domain_bits=6, lo=20, hi=32, material=FIXED_COROTATED, v0=(0.3, -1.0, 0.2), jitter_seed=None):
    """A 12^3-cell block of particles on a 64^3 grid (13 824 particles): seconds on the CPU oracle."""
    dx = 1.0 / (1 << domain_bits)
    pos = samplers.uniform_box(dx, (lo,) * 3, (hi,) * 3)
    if jitter_seed is not None:
        pos = samplers.jitter(pos, dx, 0.2, jitter_seed)
    return dict(domain_bits=domain_bits, models=[dict(material=material, pos=pos, v0=v0)])


def dense_cube(domain_bits=6, lo=20, hi=28, per_axis=3, material=FIXED_COROTATED, v0=(0.4, -0.8, 0.3)):
    """per_axis^3 particles per cell (27 -> 1728 per 4^3 block): a particle block needs several 512-particle passes of g2p2g."""
    dx = 1.0 / (1 << domain_bits)
    cells = np.arange(lo, hi, dtype=np.float64)
    sub = (np.arange(per_axis, dtype=np.float64) - (per_axis - 1) / 2) / per_axis  # offsets inside round(p / dx) == cell
    ax = (cells[:, None] + sub[None, :]).ravel() * dx
    pos = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    return dict(domain_bits=domain_bits, models=[dict(material=material, pos=pos, v0=v0)])


def two_spheres(domain_bits=8, radius=0.1645, centers=((0.30, 0.5, 0.5), (0.70, 0.5, 0.5)), speed=1.0, material=FIXED_COROTATED):
    """Configs 2 / 2b: two spheres flying at each other; radius 0.1645 -> 42.1 cells at 256^3, 84.2 at 512^3."""
    dx = 1.0 / (1 << domain_bits)
    return dict(domain_bits=domain_bits, models=[
        dict(material=material, pos=samplers.sphere(dx, centers[0], radius), v0=(speed, 0.0, 0.0)),
        dict(material=material, pos=samplers.sphere(dx, centers[1], radius), v0=(-speed, 0.0, 0.0)),
    ])


def sand_column(domain_bits=9, size=(100, 250, 100), base_y=8, material=SAND):
    """Config 3: column of size[0] x size[1] x size[2] cells x 8 particles resting just above the 2-block wall (512^3: 20 M)."""
    dx = 1.0 / (1 << domain_bits)
    n = 1 << domain_bits
    x0, z0 = (n - size[0]) // 2, (n - size[2]) // 2
    pos = samplers.uniform_box(dx, (x0, base_y, z0), (x0 + size[0], base_y + size[1], z0 + size[2]))
    return dict(domain_bits=domain_bits, models=[dict(material=material, pos=pos, v0=(0.0, 0.0, 0.0))])


def fluid_dam(domain_bits=10, size=(200, 125, 200), base=(16, 16, 16)):
    """Config 4: weakly-compressible dam of size cells x 8 particles in a corner of the domain (1024^3: 40 M)."""
    dx = 1.0 / (1 << domain_bits)
    pos = samplers.uniform_box(dx, base, tuple(b + s for b, s in zip(base, size)))
    return dict(domain_bits=domain_bits, models=[dict(material=J_FLUID, pos=pos, v0=(0.0, 0.0, 0.0))])


def two_cubes_colliding(domain_bits=6, material=FIXED_COROTATED):
    """Two 8^3-cell cubes about to touch (exercises block activation / deactivation and multi-model grids)."""
    dx = 1.0 / (1 << domain_bits)
    a = samplers.uniform_box(dx, (18, 24, 24), (26, 32, 32))
    b = samplers.uniform_box(dx, (28, 25, 25), (36, 33, 33))
    return dict(domain_bits=domain_bits, models=[dict(material=material, pos=a, v0=(2.0, 0.0, 0.0)), dict(material=material, pos=b, v0=(-2.0, 0.0, 0.0))])


def apply_material(sim, model_id, material, dx, is_oracle):
    """Material parameters used by the test scenes: volume = dx^3/8 instead of the 10x default."""
    vol = dx ** 3 / 8.0
    if material == FIXED_COROTATED:
        (sim.update_fr_parameters(model_id, 1e3, vol, 5e3, 0.4) if is_oracle else sim.update_fr_parameters(1e3, vol, 5e3, 0.4, model=model_id))
    elif material == SAND:
        (sim.update_sand_parameters(model_id, 1e3, vol, 5e3, 0.4) if is_oracle else sim.update_sand_parameters(1e3, vol, 5e3, 0.4, model=model_id))
    elif material == J_FLUID:
        (sim.update_j_fluid_parameters(model_id, 1e3, vol, 4e4, 7.15, 0.01) if is_oracle else sim.update_j_fluid_parameters(1e3, vol, 4e4, 7.15, 0.01, model=model_id))
    elif material == NACC:
        (sim.update_nacc_parameters(model_id, 1e3, vol, 5e3, 0.4, 0.5, 0.8) if is_oracle else sim.update_nacc_parameters(1e3, vol, 5e3, 0.4, 0.5, 0.8, model=model_id))


def build_oracle(ob, scene, dt=1e-4, max_blocks=4000, max_ppc=128, threads=1):
    cfg = ob.make_config(domain_bits=scene["domain_bits"], max_ppc=max_ppc)
    sim = ob.OracleSim(cfg, dt, max_blocks, threads=threads)
    dx = 1.0 / (1 << scene["domain_bits"])
    for m in scene["models"]:
        mid = sim.init_model(m["material"], m["pos"], m["v0"])
        apply_material(sim, mid, m["material"], dx, True)
    sim.initial_setup()
    return sim


def build_engine(scene, dt=1e-4, max_blocks=4000, max_ppc=128, use_graph=True, fps=0, **kw):
    import claymore_b200 as cb
    cfg = cb.Config(domain_bits=scene["domain_bits"], max_ppc=max_ppc)
    sim = cb.GmpmSimulator(dt=dt, fps=fps, config=cfg, max_blocks=max_blocks, use_graph=use_graph, **kw)
    dx = 1.0 / (1 << scene["domain_bits"])
    for m in scene["models"]:
        mid = sim.init_model(m["material"], m["pos"], m["v0"])
        apply_material(sim, mid, m["material"], dx, False)
    sim.initial_setup()
    return sim


# ---- order-free comparisons -------------------------------------------------------------------------
def key_hash(keys):
    keys = np.asarray(keys, dtype=np.int64)
    return (keys[:, 0] << 40) | (keys[:, 1] << 20) | keys[:, 2]


def grid_by_key(keys, grid):
    """dict-free alignment: returns (sorted hashes, grid blocks in that order) for the first len(grid) keys."""
    h = key_hash(keys[: len(grid)])
    o = np.argsort(h)
    return h[o], grid[o]


def match_particles(a, b, tol):
    """Bijective nearest-neighbour match of two particle sets by position; returns index array into b."""
    from scipy.spatial import cKDTree
    d, idx = cKDTree(b[:, :3]).query(a[:, :3], k=1)
    assert d.max() <= tol, f"unmatched particle: nearest distance {d.max():.3e} > {tol:.3e}"
    assert len(np.unique(idx)) == len(a), "particle match is not a bijection"
    return idx
