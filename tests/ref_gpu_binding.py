"""ctypes binding of oracle/_ref/libclaymore_ref_gpu_d<bits>.so: the REFERENCE's own GMPM kernels compiled for sm_100a
behind a minimal host loop (oracle/ref_gpu_driver.cu).  TEST INFRASTRUCTURE ONLY; needs a GPU."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHANNELS = {0: 4, 1: 12, 2: 13}


def lib_path(bits):
    return os.path.join(ROOT, "oracle", "_ref", f"libclaymore_ref_gpu_d{bits}.so")


def available(bits):
    return os.path.exists(lib_path(bits))


class RefGpuSim:
    def __init__(self, bits, material, dt_default=1e-4):
        self.L = C.CDLL(lib_path(bits))
        L = self.L
        L.refgpu_create.restype = C.c_void_p
        L.refgpu_create.argtypes = [C.c_int, C.c_float]
        L.refgpu_init_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.refgpu_setup.argtypes = [C.c_void_p]
        L.refgpu_step.argtypes = [C.c_void_p, C.c_int]
        L.refgpu_time_steps.restype = C.c_double
        L.refgpu_time_steps.argtypes = [C.c_void_p, C.c_int]
        L.refgpu_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.refgpu_keys.argtypes = [C.c_void_p, C.c_void_p]
        L.refgpu_grid.argtypes = [C.c_void_p, C.c_void_p]
        L.refgpu_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.refgpu_dt.restype = C.c_float
        L.refgpu_dt.argtypes = [C.c_void_p]
        L.refgpu_destroy.argtypes = [C.c_void_p]
        if hasattr(L, "refgpu_swap"):
            L.refgpu_swap.argtypes = [C.c_char_p, C.c_int]
        assert L.refgpu_domain_bits() == bits
        self.max_blocks = L.refgpu_max_blocks()
        self.material = material
        self.h = L.refgpu_create(material, dt_default)
        self.counts = []

    def swap(self, mask, libpath=None):
        """Route the reference kernels selected by `mask` through libclaymore_b200 via include/claymore_b200_adapter.cuh
        (bit 0 g2p2g, 1 update_grid_velocity_query_max, 2 partition / bucket / grid-carry group, 3 init kernels); 0 = all reference."""
        if libpath is None:
            import claymore_b200 as cb
            libpath = cb.lib_path()
        rc = self.L.refgpu_swap(libpath.encode(), mask)
        assert rc == 0, f"refgpu_swap failed: {rc}"

    def init_model(self, pos, v0, params):
        pos = np.ascontiguousarray(pos, np.float32)
        v0 = np.ascontiguousarray(v0, np.float32)
        p = np.ascontiguousarray(params, np.float32)
        self.counts.append(len(pos))
        return self.L.refgpu_init_model(self.h, pos.ctypes.data, len(pos), v0.ctypes.data, p.ctypes.data)

    def setup(self):
        assert self.L.refgpu_setup(self.h) == 0

    def step(self, n=1):
        assert self.L.refgpu_step(self.h, n) == 0

    def time_steps(self, n):
        return float(self.L.refgpu_time_steps(self.h, n))

    def block_counts(self):
        c = np.zeros(3, np.int32)
        self.L.refgpu_counts(self.h, c.ctypes.data)
        return tuple(int(x) for x in c)

    def active_keys(self):
        out = np.zeros((self.max_blocks, 3), np.int32)
        n = self.L.refgpu_keys(self.h, out.ctypes.data)
        return out[:n]

    def grid(self):
        _, nbc, _ = self.block_counts()
        out = np.zeros((nbc, 4, 64), np.float32)
        self.L.refgpu_grid(self.h, out.ctypes.data)
        return out

    def particle_state(self, model):
        nch = CHANNELS[self.material]
        out = np.zeros((self.counts[model], nch), np.float32)
        n = self.L.refgpu_state(self.h, model, out.ctypes.data)
        return out[:n]

    @property
    def dt(self):
        return float(self.L.refgpu_dt(self.h))

    def close(self):
        if self.h:
            self.L.refgpu_destroy(self.h)
            self.h = None


def material_params(material, dx):
    vol = dx ** 3 / 8.0
    if material == 0:
        return [1e3, vol, 4e4, 7.15, 0.01]
    return [1e3, vol, 5e3, 0.4]


def build_ref(scene, dt=1e-4, swap_mask=0):
    bits = scene["domain_bits"]
    material = scene["models"][0]["material"]
    sim = RefGpuSim(bits, material, dt)
    sim.swap(swap_mask) if (swap_mask or hasattr(sim.L, "refgpu_swap")) else None
    dx = 1.0 / (1 << bits)
    for m in scene["models"]:
        assert m["material"] == material
        sim.init_model(m["pos"], m["v0"], material_params(material, dx))
    sim.setup()
    return sim
