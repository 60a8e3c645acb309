"""Drop-in boundary test: GmpmSimulator::initial_setup / main_loop (reference Projects/GMPM/gmpm_simulator.cuh:324-580, 637-781)
re-issued call by call through the KERNEL-LEVEL C ABI (one cb200_* entry per reference kernel, caller-owned device memory, host
counters read back exactly where the reference reads them), compared with the CPU oracle.  This is the path a maintainer gets by
replacing the `compute_launch(..., kernel, ...)` lines of the reference as INTEGRATION.md shows."""
import ctypes as C

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


class DropInSim:
    def __init__(self, cb, L, torch, scene, dt_default=1e-4, max_blocks=4000, max_ppc=128):
        self.cb, self.L, self.torch = cb, L, torch
        self.cfg = cb.Config(domain_bits=scene["domain_bits"], max_ppc=max_ppc)
        self.pcfg = C.byref(self.cfg)
        self.mb = max_blocks
        self.ppb = 64 * max_ppc
        g = self.cfg.grid_size
        dev = "cuda"
        i32 = torch.int32

        def part():
            p = cb.Partition()
            t = dict(count=torch.zeros(1, dtype=i32, device=dev), index_table=torch.full((g * g * g,), -1, dtype=i32, device=dev),
                     active_keys=torch.zeros(3 * (max_blocks + 1), dtype=i32, device=dev), halo_count=torch.zeros(1, dtype=i32, device=dev),
                     halo_marks=torch.zeros(max_blocks + 1, dtype=torch.int8, device=dev), overlap_marks=torch.zeros(max_blocks + 1, dtype=i32, device=dev),
                     halo_blocks=torch.zeros(3 * (max_blocks + 1), dtype=i32, device=dev))
            for k, v in t.items():
                setattr(p, k, v.data_ptr())
            return p, t
        self.parts = [part(), part()]
        self.grids = [torch.zeros(256 * (max_blocks + 1), device=dev) for _ in range(2)]
        self.marks = torch.zeros(max_blocks + 2, dtype=i32, device=dev)
        self.dest = torch.zeros(max_blocks + 2, dtype=i32, device=dev)
        self.sources = torch.zeros(max_blocks + 2, dtype=i32, device=dev)
        self.bin_sizes = torch.zeros(max_blocks + 2, dtype=i32, device=dev)
        self.max_vel = torch.zeros(1, device=dev)
        self.models = []
        self.rollid = 0
        self.dt_default = dt_default
        self.dt = dt_default
        dx = self.cfg.dx
        import oracle_binding as ob
        for m in scene["models"]:
            n = len(m["pos"])
            binf = cb._capi.BIN_FLOATS[m["material"]]
            cap = n // 32 + max_blocks
            bufs = []
            for _ in range(2):
                t = dict(bins=torch.zeros(cap * binf, device=dev), cell_particle_counts=torch.zeros(64 * (max_blocks + 1), dtype=i32, device=dev),
                         particle_bucket_sizes=torch.zeros(max_blocks + 2, dtype=i32, device=dev), cellbuckets=torch.zeros(self.ppb * (max_blocks + 1), dtype=i32, device=dev),
                         blockbuckets=torch.zeros(self.ppb * (max_blocks + 1), dtype=i32, device=dev), bin_offsets=torch.zeros(max_blocks + 2, dtype=i32, device=dev))
                d = ob.default_buffer(ob.make_config(domain_bits=scene["domain_bits"]), m["material"])   # material defaults (host-side parameter table only)
                pb = cb.ParticleBuffer()
                for f, _t in cb.ParticleBuffer._fields_:
                    if f in t:
                        setattr(pb, f, t[f].data_ptr())
                    else:
                        setattr(pb, f, getattr(d, f))
                vol = dx ** 3 / 8.0
                pb.volume, pb.mass = vol, vol * 1e3
                bufs.append((pb, t))
            self.models.append(dict(material=m["material"], n=n, pos=torch.from_numpy(np.ascontiguousarray(m["pos"])).cuda(), v0=np.asarray(m["v0"], np.float32), bufs=bufs))

    def ck(self, rc):
        assert rc == 0, self.L.cb200_error_string(rc)

    def count_of(self, part):
        return int(part[1]["count"].item())

    def initial_setup(self):
        L, R, Rn = self.L, self.rollid, self.rollid ^ 1
        mv = max(float(np.linalg.norm(m["v0"])) for m in self.models)
        self.dt = min(self.dt_default, self.cfg.dx * self.cfg.cfl / mv) if mv > 0 else self.dt_default
        pn = self.parts[Rn]
        for m in self.models:
            self.ck(L.cb200_activate_blocks(self.pcfg, m["n"], m["pos"].data_ptr(), pn[0], None))
        self.pbc = self.count_of(pn)
        for m in self.models:
            pb, t = m["bufs"][R]
            self.ck(L.cb200_build_particle_cell_buckets(self.pcfg, m["n"], m["pos"].data_ptr(), pb, pn[0], None))
            t["particle_bucket_sizes"][: self.pbc + 1] = 0
            self.ck(L.cb200_cell_bucket_to_block(self.pcfg, self.pbc, t["cell_particle_counts"].data_ptr(), t["cellbuckets"].data_ptr(), t["particle_bucket_sizes"].data_ptr(), t["blockbuckets"].data_ptr(), None))
            self.ck(L.cb200_compute_bin_capacity(self.pbc + 1, t["particle_bucket_sizes"].data_ptr(), self.bin_sizes.data_ptr(), None))
            self.ck(L.cb200_exclusive_scan(self.pbc + 1, self.bin_sizes.data_ptr(), t["bin_offsets"].data_ptr(), None))
            self.ck(L.cb200_array_to_buffer(self.pcfg, self.pbc, m["pos"].data_ptr(), pb, None))
        self.ck(L.cb200_register_neighbor_blocks(self.pcfg, self.pbc, pn[0], None))
        self.nbc = self.count_of(pn)
        self.ck(L.cb200_register_exterior_blocks(self.pcfg, self.pbc, pn[0], None))
        self.ebc = self.count_of(pn)
        pr = self.parts[R]
        pr[1]["index_table"].copy_(pn[1]["index_table"])
        pr[1]["active_keys"].copy_(pn[1]["active_keys"])
        for m in self.models:
            m["bufs"][Rn][1]["bin_offsets"].copy_(m["bufs"][R][1]["bin_offsets"])
            m["bufs"][Rn][1]["particle_bucket_sizes"].copy_(m["bufs"][R][1]["particle_bucket_sizes"])
        self.ck(L.cb200_clear_grid(self.nbc, self.grids[0].data_ptr(), None))
        for m in self.models:
            self.ck(L.cb200_rasterize(self.pcfg, m["n"], m["pos"].data_ptr(), self.grids[0].data_ptr(), pr[0], m["bufs"][R][0].mass, m["v0"].ctypes.data_as(C.c_void_p), None))
            tn = m["bufs"][Rn][1]
            self.ck(L.cb200_init_adv_bucket(self.pcfg, self.pbc, tn["particle_bucket_sizes"].data_ptr(), tn["blockbuckets"].data_ptr(), None))
        self.torch.cuda.synchronize()

    def step(self):
        L, R, Rn = self.L, self.rollid, self.rollid ^ 1
        pr, pn = self.parts[R], self.parts[Rn]
        self.max_vel.zero_()
        self.ck(L.cb200_update_grid_velocity_query_max(self.pcfg, self.nbc, self.grids[0].data_ptr(), pr[0], self.dt, self.max_vel.data_ptr(), None))
        mv = float(np.sqrt(self.max_vel.item()))
        next_dt = min(self.dt_default, self.cfg.dx * self.cfg.cfl / mv) if mv > 0 else self.dt_default
        self.ck(L.cb200_clear_grid(self.nbc, self.grids[1].data_ptr(), None))
        for m in self.models:
            (pc, tc), (pnx, tnx) = m["bufs"][R], m["bufs"][Rn]
            tnx["cell_particle_counts"][: self.ebc * 64] = 0
            self.ck(L.cb200_g2p2g(self.pcfg, self.dt, next_dt, self.pbc, pc, pnx, pn[0], pr[0], self.grids[0].data_ptr(), self.grids[1].data_ptr(), None))
        for m in self.models:
            pnx, tnx = m["bufs"][Rn]
            tnx["particle_bucket_sizes"][: self.ebc + 1] = 0
            self.ck(L.cb200_cell_bucket_to_block(self.pcfg, self.ebc, tnx["cell_particle_counts"].data_ptr(), tnx["cellbuckets"].data_ptr(), tnx["particle_bucket_sizes"].data_ptr(), tnx["blockbuckets"].data_ptr(), None))
        self.marks[: self.nbc] = 0
        self.ck(L.cb200_mark_active_grid_blocks(self.nbc, self.grids[1].data_ptr(), self.marks.data_ptr(), None))
        self.sources[: self.ebc + 1] = 0
        for m in self.models:
            self.ck(L.cb200_mark_active_particle_blocks(self.ebc + 1, m["bufs"][Rn][1]["particle_bucket_sizes"].data_ptr(), self.sources.data_ptr(), None))
        self.ck(L.cb200_exclusive_scan(self.ebc + 1, self.sources.data_ptr(), self.dest.data_ptr(), None))
        new_pbc = int(self.dest[self.ebc].item())
        pn[1]["count"][0] = new_pbc
        self.ck(L.cb200_exclusive_scan_inverse(self.ebc, self.dest.data_ptr(), self.sources.data_ptr(), None))
        self.ck(L.cb200_reset_table(self.pcfg, pn[0], None))
        self.ck(L.cb200_update_partition(self.pcfg, new_pbc, self.sources.data_ptr(), pr[0], pn[0], None))
        for m in self.models:
            (pc, tc), (pnx, tnx) = m["bufs"][R], m["bufs"][Rn]
            self.ck(L.cb200_update_buckets(self.pcfg, new_pbc, self.sources.data_ptr(), pnx, pc, None))
            tc["particle_bucket_sizes"][new_pbc] = 0
            self.ck(L.cb200_compute_bin_capacity(new_pbc + 1, tc["particle_bucket_sizes"].data_ptr(), self.bin_sizes.data_ptr(), None))
            self.ck(L.cb200_exclusive_scan(new_pbc + 1, self.bin_sizes.data_ptr(), tc["bin_offsets"].data_ptr(), None))
        self.ck(L.cb200_register_neighbor_blocks(self.pcfg, new_pbc, pn[0], None))
        prev_nbc, new_nbc = self.nbc, self.count_of(pn)
        self.ck(L.cb200_clear_grid(max(self.ebc, new_nbc), self.grids[0].data_ptr(), None))
        self.ck(L.cb200_copy_selected_grid_blocks(self.pcfg, prev_nbc, pr[1]["active_keys"].data_ptr(), pn[0], self.marks.data_ptr(), self.grids[1].data_ptr(), self.grids[0].data_ptr(), None))
        self.ck(L.cb200_register_exterior_blocks(self.pcfg, new_pbc, pn[0], None))
        self.pbc, self.nbc, self.ebc = new_pbc, new_nbc, self.count_of(pn)
        self.rollid, self.dt = Rn, next_dt

    # the observation interface compare_state expects
    def block_counts(self):
        return self.pbc, self.nbc, self.ebc

    def active_keys(self):
        return self.parts[self.rollid][1]["active_keys"][: 3 * self.ebc].cpu().numpy().reshape(-1, 3)

    def grid(self):
        return self.grids[0][: 256 * self.nbc].cpu().numpy().reshape(-1, 4, 64)

    def particle_state(self, mi):
        m = self.models[mi]
        R, Rn = self.rollid, self.rollid ^ 1
        out = self.torch.zeros(m["n"] * 3, device="cuda")
        cnt = self.torch.zeros(1, dtype=self.torch.int32, device="cuda")
        self.ck(self.L.cb200_retrieve_particle_buffer(self.pcfg, self.pbc, self.parts[R][0], self.parts[Rn][0], m["bufs"][R][0], m["bufs"][Rn][0], out.data_ptr(), cnt.data_ptr(), None))
        n = int(cnt.item())
        return out[: 3 * n].cpu().numpy().reshape(-1, 3)


@pytest.mark.parametrize("case", ["fc_cube", "two_models"])
def test_reference_loop_through_kernel_level_abi(oracle, cuda_lib, case):
    import torch
    import claymore_b200 as cb
    scene = scenes.small_cube() if case == "fc_cube" else scenes.two_cubes_colliding()
    dt = 1e-4 if case == "fc_cube" else 2e-4
    osim = scenes.build_oracle(oracle, scene, dt=dt)
    dsim = DropInSim(cb, cuda_lib, torch, scene, dt_default=dt)
    dsim.initial_setup()
    for chunk in range(3):
        for _ in range(4):
            osim.step(1)
            dsim.step()
        assert dsim.block_counts() == osim.block_counts()
        ok, dk = osim.active_keys(), dsim.active_keys()
        pbc, nbc, ebc = osim.block_counts()
        for lo, hi in ((0, pbc), (pbc, nbc), (nbc, ebc)):
            assert np.array_equal(np.sort(scenes.key_hash(ok[lo:hi])), np.sort(scenes.key_hash(dk[lo:hi])))
        oh, og = scenes.grid_by_key(ok, osim.grid())
        dh, dg = scenes.grid_by_key(dk, dsim.grid())
        assert np.array_equal(oh, dh)
        assert np.allclose(dg[:, 0], og[:, 0], rtol=1e-5, atol=1e-5 * og[:, 0].max())
        assert np.abs(dg[:, 1:] - og[:, 1:]).max() <= 1e-4 * np.abs(og[:, 1:]).max()
        for mi in range(len(scene["models"])):
            so, sd = osim.retrieve(mi), dsim.particle_state(mi)
            assert len(so) == len(sd)
            scenes.match_particles(so, sd, tol=5e-6)
        assert abs(dsim.dt - osim.dt) <= 1e-9


# ---- the reference's own host loop and containers with single kernels swapped for the library's --------------------------------
@pytest.mark.parametrize("material", [scenes.FIXED_COROTATED, scenes.J_FLUID, scenes.SAND])
@pytest.mark.parametrize("mask,what", [(1, "g2p2g"), (2, "update_grid_velocity_query_max"), (4, "partition / bucket / carry group"), (8, "init kernels"), (15, "everything")])
def test_reference_loop_with_swapped_kernels(cuda_lib, material, mask, what):
    """oracle/ref_gpu_driver.cu drives the reference's REAL ParticleBuffer<M> / Partition<1> / GridBuffer instances through the
    reference's launch sequence (gmpm_simulator.cuh:324-580, 637-781); refgpu_swap routes the selected kernels through
    include/claymore_b200_adapter.cuh -> the C ABI.  The swapped run must reproduce the all-reference run: block counts and key
    sets per class bit-exact, per-cell mass 2e-5, momentum 2e-4 of the max, particle states matched by position."""
    import ref_gpu_binding as rg
    from test_oracle_cpu import compare_with_ref_gpu_golden
    if not rg.available(6):
        pytest.skip("oracle/_ref/libclaymore_ref_gpu_d6.so not built")
    if mask not in (1, 15) and material != scenes.FIXED_COROTATED:
        pytest.skip("material independent kernels: covered once")
    scene = scenes.small_cube(material=material, jitter_seed=5)
    ref = rg.build_ref(scene, swap_mask=0)
    ref.step(12)
    pbc, nbc, ebc = ref.block_counts()
    keys = ref.active_keys()
    h = scenes.key_hash(keys)
    gh, gg = scenes.grid_by_key(keys, ref.grid())
    golden = {"s12_counts": np.array([pbc, nbc, ebc]), "s12_keys_particle": np.sort(h[:pbc]), "s12_keys_neighbor": np.sort(h[pbc:nbc]),
              "s12_keys_exterior": np.sort(h[nbc:ebc]), "s12_grid_keys": gh, "s12_grid": gg, "s12_state0": ref.particle_state(0)}
    ref.close()
    swapped = rg.build_ref(scene, swap_mask=mask)
    swapped.step(12)
    compare_with_ref_gpu_golden(swapped, golden, 12, 1, f"reference loop with {what} swapped")
    swapped.close()
    rg.RefGpuSim(6, material).swap(0)   # leave the shared library in its all-reference state
