"""ctypes binding of the CPU oracle (oracle/libclaymore_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  The product package (claymore_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "libclaymore_oracle.so")

J_FLUID, FIXED_COROTATED, SAND, NACC = 0, 1, 2, 3
CHANNELS = {J_FLUID: 4, FIXED_COROTATED: 12, SAND: 13, NACC: 13}
BIN_FLOATS = {J_FLUID: 128, FIXED_COROTATED: 512, SAND: 512, NACC: 512}


class Config(C.Structure):
    _fields_ = [("domain_bits", C.c_int), ("max_ppc", C.c_int), ("boundary", C.c_int), ("gravity", C.c_float), ("cfl", C.c_float)]


def make_config(domain_bits=8, max_ppc=128, boundary=2, gravity=-9.8, cfl=0.5):
    return Config(domain_bits, max_ppc, boundary, gravity, cfl)


class ParticleBuffer(C.Structure):
    _fields_ = [
        ("material", C.c_int),
        ("bins", C.c_void_p),
        ("cell_particle_counts", C.c_void_p),
        ("particle_bucket_sizes", C.c_void_p),
        ("cellbuckets", C.c_void_p),
        ("blockbuckets", C.c_void_p),
        ("bin_offsets", C.c_void_p),
        ("rho", C.c_float), ("volume", C.c_float), ("mass", C.c_float),
        ("bulk", C.c_float), ("gamma", C.c_float), ("viscosity", C.c_float),
        ("lambda_", C.c_float), ("mu", C.c_float),
        ("cohesion", C.c_float), ("beta", C.c_float), ("yield_surface", C.c_float),
        ("volume_correction", C.c_int),
        ("bm", C.c_float), ("xi", C.c_float), ("msqr", C.c_float),
        ("hardening_on", C.c_int),
    ]


class Partition(C.Structure):
    _fields_ = [
        ("count", C.c_void_p), ("index_table", C.c_void_p), ("active_keys", C.c_void_p),
        ("halo_count", C.c_void_p), ("halo_marks", C.c_void_p), ("overlap_marks", C.c_void_p), ("halo_blocks", C.c_void_p),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_ORACLE_DIR, "claymore_oracle.c")):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.orc_sim_create.restype = C.c_void_p
        L.orc_sim_create.argtypes = [C.POINTER(Config), C.c_float, C.c_int]
        L.orc_sim_destroy.argtypes = [C.c_void_p]
        L.orc_sim_init_model.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sim_initial_setup.argtypes = [C.c_void_p]
        L.orc_sim_step.argtypes = [C.c_void_p, C.c_float]
        L.orc_sim_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_sim_dt.restype = C.c_float
        L.orc_sim_dt.argtypes = [C.c_void_p]
        L.orc_sim_max_vel.restype = C.c_float
        L.orc_sim_max_vel.argtypes = [C.c_void_p]
        L.orc_sim_retrieve.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sim_active_keys.restype = C.c_void_p
        L.orc_sim_active_keys.argtypes = [C.c_void_p]
        L.orc_sim_grid.restype = C.c_void_p
        L.orc_sim_grid.argtypes = [C.c_void_p]
        L.orc_sim_particle_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sim_update_fr_parameters.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
        L.orc_sim_update_sand_parameters.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
        L.orc_sim_update_j_fluid_parameters.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 5
        L.orc_sim_update_nacc_parameters.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 6
        L.orc_svd3.argtypes = [C.c_void_p] * 4
        L.orc_compute_stress.argtypes = [C.c_int, C.POINTER(ParticleBuffer), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bspline_weight.argtypes = [C.POINTER(Config), C.c_float, C.c_void_p]
        L.orc_g2p2g.argtypes = [C.POINTER(Config), C.c_float, C.c_float, C.c_int, ParticleBuffer, ParticleBuffer, Partition, Partition, C.c_void_p, C.c_void_p]
        L.orc_update_grid_velocity_query_max.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, Partition, C.c_float, C.c_void_p]
        L.orc_mark_overlapping_blocks.argtypes = [C.POINTER(Config), C.c_int, C.c_int, C.c_void_p, Partition, C.c_void_p, C.c_void_p]
        L.orc_collect_blockids_for_halo_reduction.argtypes = [C.POINTER(Config), C.c_int, Partition]
        L.orc_collect_grid_blocks.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.c_void_p, Partition, C.c_void_p]
        L.orc_reduce_grid_blocks.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.c_void_p, Partition, C.c_void_p]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_sim_get_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ParticleBuffer)]
        L.orc_sim_get_partition.argtypes = [C.c_void_p, C.c_int, C.POINTER(Partition)]
        L.orc_sim_get_grid.restype = C.c_void_p
        L.orc_sim_get_grid.argtypes = [C.c_void_p, C.c_int]
        L.orc_sim_bin_capacity.restype = C.c_long
        L.orc_sim_bin_capacity.argtypes = [C.c_void_p, C.c_int]
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def svd3(F):
    """F: (9,) column-major float32 -> U(9), S(3), V(9)."""
    F = np.ascontiguousarray(F, dtype=np.float32)
    U = np.zeros(9, np.float32)
    S = np.zeros(3, np.float32)
    V = np.zeros(9, np.float32)
    lib().orc_svd3(ptr(F), ptr(U), ptr(S), ptr(V))
    return U, S, V


def default_buffer(cfg, material):
    """Material defaults of ParticleBuffer<M> (reference particle_buffer.cuh:141-264)."""
    cells = float(1 << cfg.domain_bits)
    E, nu = 5e3, 0.4
    pb = ParticleBuffer()
    pb.material = material
    pb.rho = 1e3
    pb.mass = 1e3 / cells / cells / cells / 8.0
    pb.volume = (10.0 if material in (FIXED_COROTATED, SAND) else 1.0) / cells / cells / cells / 8.0
    pb.bulk, pb.gamma, pb.viscosity = 4e4, 7.15, 0.01
    pb.lambda_ = E * nu / ((1 + nu) * (1 - 2 * nu))
    pb.mu = E / (2 * (1 + nu))
    pb.cohesion = 0.0
    pb.beta = 0.5 if material == NACC else 1.0
    pb.yield_surface = 0.816496580927726 * 2.0 * 0.5 / (3.0 - 0.5)
    pb.volume_correction = 1
    pb.bm = 2.0 / 3.0 * (E / (2 * (1 + nu))) + (E * nu / ((1 + nu) * (1 - 2 * nu)))
    pb.xi = 0.8
    pb.msqr = 3.423772074299613
    pb.hardening_on = 1
    return pb


def compute_stress(material, pb, F, log_jp=0.0):
    F = np.array(F, dtype=np.float32).copy()
    PF = np.zeros(9, np.float32)
    lj = np.array([log_jp], np.float32)
    lib().orc_compute_stress(material, C.byref(pb), ptr(F), ptr(PF), ptr(lj))
    return F, PF, float(lj[0])


class OracleSim:
    """GmpmSimulator restated on the CPU (see oracle/claymore_oracle.c)."""

    def __init__(self, cfg, dt_default=1e-4, max_blocks=20000, threads=1):
        self.cfg = cfg
        self.max_blocks = max_blocks
        self.L = lib()
        self.L.orc_set_num_threads(threads)
        self.h = self.L.orc_sim_create(C.byref(cfg), dt_default, max_blocks)
        self.materials = []
        self.counts = []

    def close(self):
        if self.h:
            self.L.orc_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def init_model(self, material, pos, v0):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        v0 = np.ascontiguousarray(v0, dtype=np.float32)
        self.materials.append(material)
        self.counts.append(len(pos))
        return self.L.orc_sim_init_model(self.h, material, ptr(pos), len(pos), ptr(v0))

    def update_fr_parameters(self, model, rho, vol, ym, pr):
        self.L.orc_sim_update_fr_parameters(self.h, model, rho, vol, ym, pr)

    def update_sand_parameters(self, model, rho, vol, ym, pr):
        self.L.orc_sim_update_sand_parameters(self.h, model, rho, vol, ym, pr)

    def update_j_fluid_parameters(self, model, rho, vol, bulk, gamma, visc):
        self.L.orc_sim_update_j_fluid_parameters(self.h, model, rho, vol, bulk, gamma, visc)

    def update_nacc_parameters(self, model, rho, vol, ym, pr, beta, xi):
        self.L.orc_sim_update_nacc_parameters(self.h, model, rho, vol, ym, pr, beta, xi)

    def initial_setup(self):
        self.L.orc_sim_initial_setup(self.h)

    def step(self, n=1, time_left=1e30):
        for _ in range(n):
            self.L.orc_sim_step(self.h, time_left)

    def block_counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_sim_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    @property
    def dt(self):
        return self.L.orc_sim_dt(self.h)

    @property
    def max_vel(self):
        return self.L.orc_sim_max_vel(self.h)

    def retrieve(self, model):
        out = np.zeros((self.counts[model], 3), np.float32)
        n = self.L.orc_sim_retrieve(self.h, model, ptr(out))
        return out[:n]

    def particle_state(self, model):
        nch = CHANNELS[self.materials[model]]
        out = np.zeros((self.counts[model], nch), np.float32)
        n = self.L.orc_sim_particle_state(self.h, model, ptr(out))
        return out[:n]

    def active_keys(self):
        _, _, ebc = self.block_counts()
        p = self.L.orc_sim_active_keys(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(ebc, 3)).copy()

    def grid(self):
        _, nbc, _ = self.block_counts()
        p = self.L.orc_sim_grid(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(nbc, 4, 64)).copy()


    # ---- raw containers (numpy views of the oracle's host memory) for kernel-level differential tests ----
    def buffer(self, model, which):
        pb = ParticleBuffer()
        self.L.orc_sim_get_buffer(self.h, model, which, C.byref(pb))
        return pb

    def partition(self, which):
        p = Partition()
        self.L.orc_sim_get_partition(self.h, which, C.byref(p))
        return p

    def buffer_arrays(self, model, which):
        """dict of numpy views: bins, cell_particle_counts, particle_bucket_sizes, cellbuckets, blockbuckets, bin_offsets."""
        pb = self.buffer(model, which)
        mb, ppb = self.max_blocks, 64 * self.cfg.max_ppc
        cap = self.L.orc_sim_bin_capacity(self.h, model)
        bf = BIN_FLOATS[self.materials[model]]

        def view(ptr_, ctype, n):
            return np.ctypeslib.as_array(C.cast(ptr_, C.POINTER(ctype)), shape=(n,))
        return {
            "struct": pb,
            "bins": view(pb.bins, C.c_float, cap * bf),
            "cell_particle_counts": view(pb.cell_particle_counts, C.c_int, mb * 64),
            "particle_bucket_sizes": view(pb.particle_bucket_sizes, C.c_int, mb + 1),
            "cellbuckets": view(pb.cellbuckets, C.c_int, mb * ppb),
            "blockbuckets": view(pb.blockbuckets, C.c_int, mb * ppb),
            "bin_offsets": view(pb.bin_offsets, C.c_int, mb + 1),
        }

    def partition_arrays(self, which):
        p = self.partition(which)
        g = 1 << (self.cfg.domain_bits - 2)

        def view(ptr_, ctype, n):
            return np.ctypeslib.as_array(C.cast(ptr_, C.POINTER(ctype)), shape=(n,))
        return {"struct": p, "count": view(p.count, C.c_int, 1), "index_table": view(p.index_table, C.c_int, g * g * g),
                "active_keys": view(p.active_keys, C.c_int, self.max_blocks * 3)}

    def grid_array(self, which):
        p = self.L.orc_sim_get_grid(self.h, which)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.max_blocks * 256,))
