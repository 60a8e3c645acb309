"""Synthetic scenes of the BASELINE configs and the small parity scenes.  Pure numpy.

A scene is ``dict(domain_bits=..., models=[dict(material=..., pos=float32[n,3], v0=(3,))])``: the arguments of
GmpmSimulator::init_model (Projects/GMPM/gmpm_simulator.cuh:168-209), one entry per model.
"""
import numpy as np

from . import samplers
from ._capi import Config, FIXED_COROTATED, J_FLUID, NACC, SAND

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


def mixed_materials(domain_bits=10, edge=None, gap=None, base=None):
    """Config 5: eight boxes (4 along x, 2 along z) of edge^3 cells x 8 particles, 20-cell gaps, resting just above the floor wall:
    fluid x4, sand x2, fixed-corotated x2, interleaved so that every x-slab of a static partition holds several materials and
    every body has at least two neighbours once the fluids spread.  1024^3: edge 116 -> 8 x 12.49 M = 99.9 M particles.
    One model per MATERIAL (the per-model bucket arrays are sized by the block capacity)."""
    n = 1 << domain_bits
    edge = edge if edge is not None else (116 * n) // 1024
    gap = gap if gap is not None else max(4, (20 * n) // 1024)
    base = base if base is not None else max(10, (16 * n) // 1024)
    dx = 1.0 / n
    x0 = (n - (4 * edge + 3 * gap)) // 2
    z0 = (n - (2 * edge + gap)) // 2
    layout = [[J_FLUID, SAND, J_FLUID, FIXED_COROTATED], [FIXED_COROTATED, J_FLUID, SAND, J_FLUID]]
    per_mat = {J_FLUID: [], SAND: [], FIXED_COROTATED: []}
    for kz in range(2):
        for kx in range(4):
            lo = (x0 + kx * (edge + gap), base, z0 + kz * (edge + gap))
            per_mat[layout[kz][kx]].append(samplers.uniform_box(dx, lo, tuple(c + edge for c in lo)))
    return dict(domain_bits=domain_bits, models=[dict(material=mat, pos=np.concatenate(per_mat[mat], 0), v0=(0.0, 0.0, 0.0)) for mat in (J_FLUID, SAND, FIXED_COROTATED)])


def two_cubes_colliding(domain_bits=6, material=FIXED_COROTATED):
    """Two 8^3-cell cubes about to touch (exercises block activation / deactivation and multi-model grids)."""
    dx = 1.0 / (1 << domain_bits)
    a = samplers.uniform_box(dx, (18, 24, 24), (26, 32, 32))
    b = samplers.uniform_box(dx, (28, 25, 25), (36, 33, 33))
    return dict(domain_bits=domain_bits, models=[dict(material=material, pos=a, v0=(2.0, 0.0, 0.0)), dict(material=material, pos=b, v0=(-2.0, 0.0, 0.0))])


def material_parameters(material, dx):
    """(setter name, arguments) of the test scenes' materials: volume = dx^3/8 instead of the 10x default."""
    vol = dx ** 3 / 8.0
    if material == FIXED_COROTATED:
        return "update_fr_parameters", (1e3, vol, 5e3, 0.4)
    if material == SAND:
        return "update_sand_parameters", (1e3, vol, 5e3, 0.4)
    if material == J_FLUID:
        return "update_j_fluid_parameters", (1e3, vol, 4e4, 7.15, 0.01)
    if material == NACC:
        return "update_nacc_parameters", (1e3, vol, 5e3, 0.4, 0.5, 0.8)
    raise ValueError(material)


def apply_material(sim, model_id, material, dx):
    name, args = material_parameters(material, dx)
    getattr(sim, name)(*args, model=model_id)


def build_engine(scene, dt=1e-4, max_blocks=4000, max_ppc=128, use_graph=True, fps=0, cfl=0.5, **kw):
    from .simulator import GmpmSimulator
    cfg = Config(domain_bits=scene["domain_bits"], max_ppc=max_ppc, cfl=cfl)
    sim = GmpmSimulator(dt=dt, fps=fps, config=cfg, max_blocks=max_blocks, use_graph=use_graph, **kw)
    dx = 1.0 / (1 << scene["domain_bits"])
    for m in scene["models"]:
        mid = sim.init_model(m["material"], m["pos"], m["v0"])
        apply_material(sim, mid, m["material"], dx)
    sim.initial_setup()
    return sim


def n_particles(scene):
    return sum(len(m["pos"]) for m in scene["models"])


def max_blocks_for(scene, factor=2.5):
    """Block capacity (the reference's G_MAX_ACTIVE_BLOCK) for a scene: particle blocks at 8 ppc x head-room for the shell."""
    return int(max(4000, n_particles(scene) / 512 * factor))


# ---- named workloads of bench.py / the parity tests (BASELINE.json configs) ---------------------------------------------
def workload(name):
    """(scene, label) of a named workload."""
    if name == "spheres5m":
        return two_spheres(domain_bits=8), "GMPM two elastic spheres (fixed-corotated), 256^3 grid, 5M particles"
    if name == "spheres40m":
        return two_spheres(domain_bits=9), "GMPM two elastic spheres (fixed-corotated), 512^3 grid, 40M particles"
    if name == "sand20m":
        return sand_column(), "GMPM sand column collapse (Drucker-Prager), 512^3 grid, 20M particles"
    if name == "sand2m":
        return sand_column(domain_bits=8, size=(50, 100, 50)), "sand column (Drucker-Prager), 256^3 grid, 2M particles"
    if name == "fluid40m":
        return fluid_dam(), "weakly-compressible fluid dam break, 1024^3 grid, 40M particles"
    if name == "fluid5m":
        return fluid_dam(domain_bits=9, size=(100, 62, 100)), "weakly-compressible fluid dam, 512^3 grid, 5M particles"
    if name == "mixed100m":
        return mixed_materials(), "MGSP mixed materials (4 fluid + 2 sand + 2 fixed-corotated bodies), 1024^3 grid, 100M particles"
    if name == "mixed12m":
        return mixed_materials(domain_bits=9), "mixed materials (4 fluid + 2 sand + 2 fixed-corotated bodies), 512^3 grid, 12.5M particles"
    if name == "cube140k":
        return jelly_cube(), "jelly cube (fixed-corotated), 128^3 grid, 140608 particles"
    if name == "spheres640k":
        return two_spheres(domain_bits=8, radius=0.1645 / 2), "two elastic spheres (fixed-corotated), 256^3 grid, r=21 cells, 0.64M particles"
    raise SystemExit(f"unknown workload {name}")


# ---- order-free comparisons -------------------------------------------------------------------------
def key_hash(keys):
    keys = np.asarray(keys, dtype=np.int64)
    return (keys[:, 0] << 40) | (keys[:, 1] << 20) | keys[:, 2]


def grid_by_key(keys, grid):
    """dict-free alignment: returns (sorted hashes, grid blocks in that order) for the first len(grid) keys."""
    h = key_hash(keys[: len(grid)])
    o = np.argsort(h)
    return h[o], grid[o]
