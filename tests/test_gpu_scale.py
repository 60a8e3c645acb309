"""GPU parity at the scales that are benchmarked, device math unit tests, capacity growth / overflow, multi-rank MGSP.

  * engine vs the LIVE reference kernels (oracle/_ref/libclaymore_ref_gpu_d<bits>.so: the reference's own sources built for
    sm_100a, travels with the snapshot) on BASELINE config 2 (5 M particles, 256^3) after 1 / 20 / 100 sub-steps and on a 2 M-particle
    sand column / 1.3 M-particle fluid dam (512^3): block counts and key sets per class bit-exact, per-cell mass 2e-5 of the max,
    momentum 2e-4 of the max, totals 1e-6, particle count, a 200 k-particle sample matched by position;
  * the device SVD / constitutive models on the 4000 golden vectors recorded from the reference's own svd.cuh /
    constitutive_models.cuh (tests/golden/ref_math_golden.npz), incl. large strain, inverted and nearly singular F;
  * block-capacity overflow sets the error bit without touching memory outside the containers; in-place growth continues a run
    that starts under-provisioned and stays identical to a run that was provisioned generously;
  * MGSP with FOUR ranks in a 2 x 2 split (grid blocks shared by four ranks) in one process, and world-size 2 / 4 multi-process
    runs over CUDA IPC (skipped when the box has fewer GPUs) against the single-GPU engine.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_math_golden.npz")


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


# ---------------------------------------------------------------------------------------------------------------------------
# engine vs live reference kernels at benchmark scale
# ---------------------------------------------------------------------------------------------------------------------------
def _compare_with_live_reference(ref, esim, nmodels, label, sample=200_000, pos_tol=5e-6, f_tol=5e-4, mass_tol=2e-5, mom_tol=2e-4):
    st = esim.stats()
    assert st.error == 0, (label, st.error)
    pbc, nbc, ebc = ref.block_counts()
    assert (st.particle_block_count, st.neighbor_block_count, st.exterior_block_count) == (pbc, nbc, ebc), label
    rk, ek = ref.active_keys(), esim.active_keys()
    for lo, hi in ((0, pbc), (pbc, nbc), (nbc, ebc)):  # per class: particle / neighbour / exterior -- bit-exact as sets
        assert np.array_equal(np.sort(scenes.key_hash(rk[lo:hi])), np.sort(scenes.key_hash(ek[lo:hi]))), f"{label}: key set of class [{lo},{hi})"
    rh, rg = scenes.grid_by_key(rk, ref.grid())
    eh, eg = scenes.grid_by_key(ek, esim.grid())
    assert np.array_equal(rh, eh)
    mass_r, mass_e = rg[:, 0], eg[:, 0]
    mscale = float(mass_r.max())
    assert np.abs(mass_e - mass_r).max() <= mass_tol * mscale, f"{label}: cell mass {np.abs(mass_e - mass_r).max() / mscale:.3e} of max"
    pscale = float(np.abs(rg[:, 1:]).max())
    assert np.abs(eg[:, 1:] - rg[:, 1:]).max() <= mom_tol * pscale, f"{label}: cell momentum {np.abs(eg[:, 1:] - rg[:, 1:]).max() / pscale:.3e} of max"
    tm_r, tm_e = mass_r.sum(dtype=np.float64), mass_e.sum(dtype=np.float64)
    assert abs(tm_e - tm_r) <= 1e-6 * tm_r, label
    tp_r, tp_e = rg[:, 1:].sum(axis=(0, 2), dtype=np.float64), eg[:, 1:].sum(axis=(0, 2), dtype=np.float64)
    tot_tol = 1e-5 if mom_tol <= 2e-4 else 5e-5   # branch-point noise (see the sand test) also shows in the total (measured 1.1e-5)
    assert np.abs(tp_e - tp_r).max() <= tot_tol * np.abs(rg[:, 1:]).sum(dtype=np.float64) / 3 + 1e-12, (label, tp_e, tp_r)
    rng = np.random.default_rng(1)
    for m in range(nmodels):
        sr, se = ref.particle_state(m), esim.particle_state(m)
        assert len(sr) == len(se), f"{label}: particle count of model {m}"
        pick = rng.choice(len(sr), size=min(sample, len(sr)), replace=False)
        from scipy.spatial import cKDTree
        d, idx = cKDTree(se[:, :3]).query(sr[pick, :3], k=1)
        assert d.max() <= pos_tol, f"{label}: position {d.max():.3e}"
        if sr.shape[1] > 3:
            err = np.abs(se[idx][:, 3:] - sr[pick][:, 3:]).max()
            assert err <= f_tol, f"{label}: particle channels {err:.3e}"


@pytest.mark.timeout(900)
def test_config2_5m_vs_live_reference(cuda_lib):
    """BASELINE configs[1]: two fixed-corotated spheres, 256^3 grid, 5 M particles -- ~11.7 k particle blocks, ~30 k exterior blocks."""
    import ref_gpu_binding as rg
    if not rg.available(8):
        pytest.skip("oracle/_ref/libclaymore_ref_gpu_d8.so not built")
    scene, _ = scenes.workload("spheres5m")
    ref = rg.build_ref(scene)
    esim = scenes.build_engine(scene, max_blocks=scenes.max_blocks_for(scene))
    done = 0
    for cp in (1, 20, 100):
        ref.step(cp - done)
        esim.step(cp - done)
        done = cp
        # fast-math reference, different summation orders: particle channels drift apart slowly over 100 sub-steps
        _compare_with_live_reference(ref, esim, 2, f"spheres5m step {cp}", pos_tol=3e-6 if cp < 100 else 1e-5, f_tol=2e-4 if cp < 100 else 1e-3)
    assert sum(len(esim.retrieve(i)) for i in range(2)) == scenes.n_particles(scene)
    ref.close()
    esim.close()


def _grid_deviation(a, b):
    """(max |cell mass a - b| / max mass, max |cell momentum a - b| / max |momentum|) of two simulators' grids aligned by block key."""
    ha, ga = scenes.grid_by_key(a.active_keys(), a.grid())
    hb, gb = scenes.grid_by_key(b.active_keys(), b.grid())
    assert np.array_equal(ha, hb)
    return float(np.abs(ga[:, 0] - gb[:, 0]).max() / ga[:, 0].max()), float(np.abs(ga[:, 1:] - gb[:, 1:]).max() / np.abs(ga[:, 1:]).max())


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("name", ["sand2m_512", "fluid1m_512"])
def test_sand_and_fluid_1m_vs_live_reference_and_oracle(oracle, cuda_lib, name):
    """>= 1 M particles on the 512^3 grid, 20 sub-steps, THREE-WAY: the engine, the LIVE reference kernels (fast-math build of the
    reference's own sources) and the oracle (exact-arithmetic restatement, pinned bitwise to the reference's math).
    Structure (block counts, key sets per class), totals and particle counts must agree exactly / to 1e-6 with both.
    Per-cell fields: after one sub-step 2e-5 (mass) / 2e-4 (momentum) of the max against both.  After 20 sub-steps the sand column has
    particles ON the Drucker-Prager yield surface; which side of a branch a particle takes depends on the last bits of its singular
    values, so even the reference and its own exact restatement -- which share ONE SVD algorithm, operation for operation -- differ
    there (measured: 5.4e-6 of the max cell mass, 5.1e-4 of the max cell momentum).  The engine's SVD is an own implementation of the
    same scheme (different rounding in the last bits), so it sits a little further out: measured 2.3e-5 / 2.1e-3 against the reference,
    1.8e-5 / 1.9e-3 against the oracle -- a velocity difference of 4e-5 m/s on cells moving at 2e-2 m/s.  The bound is relative to
    that spread: 5x the reference-vs-oracle distance (floor 2e-5 mass / 2e-4 momentum); the fluid dam, which has no yield branch,
    stays at the floor (measured 5e-7 / 2e-6)."""
    import ref_gpu_binding as rg
    if not rg.available(9):
        pytest.skip("oracle/_ref/libclaymore_ref_gpu_d9.so not built")
    sand = name.startswith("sand")
    scene = scenes.sand_column(domain_bits=9, size=(50, 100, 50)) if sand else scenes.fluid_dam(domain_bits=9, size=(64, 40, 64), base=(12, 12, 12))
    assert scenes.n_particles(scene) >= 1_000_000
    mb = scenes.max_blocks_for(scene, 4.0)
    ref = rg.build_ref(scene)
    esim = scenes.build_engine(scene, max_blocks=mb)
    osim = scenes.build_oracle(oracle, scene, max_blocks=mb, threads=min(os.cpu_count() or 1, 32))
    done = 0
    for cp in (1, 20):
        ref.step(cp - done)
        esim.step(cp - done)
        osim.step(cp - done)
        done = cp
        m_ro, p_ro = _grid_deviation(ref, osim)
        mass_tol, mom_tol = max(2e-5, 5 * m_ro), max(2e-4, 5 * p_ro)
        print(f"{name} step {cp}: reference vs oracle mass {m_ro:.2e} momentum {p_ro:.2e} -> engine bounds {mass_tol:.2e} / {mom_tol:.2e}; "
              f"engine vs reference {_grid_deviation(ref, esim)}, engine vs oracle {_grid_deviation(osim, esim)}")
        _compare_with_live_reference(ref, esim, 1, f"{name} step {cp} vs live reference", f_tol=1e-3, mass_tol=mass_tol, mom_tol=mom_tol)
        _compare_with_live_reference(osim, esim, 1, f"{name} step {cp} vs oracle", f_tol=1e-3, mass_tol=mass_tol, mom_tol=mom_tol)
    ref.close()
    esim.close()
    osim.close()


# ---------------------------------------------------------------------------------------------------------------------------
# device math on the reference's golden vectors
# ---------------------------------------------------------------------------------------------------------------------------
def _default_buffer(cb, material, bits=8):
    cfg = cb.Config(domain_bits=bits)
    pb = cb.ParticleBuffer()
    cb.lib().cb200_default_material(C.byref(cfg), material, C.byref(pb))
    return pb


def test_device_svd_on_reference_golden(cuda_lib):
    """svd3 (math3.cuh) is an own implementation of the reference's scheme (math::svd, svd.cuh:28-1124: four cyclic Jacobi sweeps
    with approximate Givens angles, then a Givens QR), so it is held to that scheme's accuracy class, measured on the reference's own
    output for the same 4000 vectors (reconstruction error relative to max(1, |F|): median 4.4e-7, 99.9 % below 2.2e-4, worst 1.7e-3;
    singular values within 4e-5 of the exact ones):
      reconstruction: median <= 1e-6, 99.9 % <= 5e-4, worst <= 5e-3;  U, V orthogonal to 5e-6 with det +1;  |S| sorted;
      singular values within 1e-4 max(1, |F|) of the exact ones (float64 LAPACK) and of the reference's;  sign of det F on the last value."""
    torch = _torch()
    import claymore_b200 as cb
    g = np.load(GOLDEN)
    F = g["F"].astype(np.float32)
    n = len(F)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tF, tU, tS, tV = d(F), torch.zeros(n, 9, device="cuda"), torch.zeros(n, 3, device="cuda"), torch.zeros(n, 9, device="cuda")
    cb._capi.check(cb.lib().cb200_test_svd3(n, tF.data_ptr(), tU.data_ptr(), tS.data_ptr(), tV.data_ptr(), None), "test_svd3")
    torch.cuda.synchronize()
    U = tU.cpu().numpy().reshape(n, 3, 3).transpose(0, 2, 1).astype(np.float64)   # column-major -> [r][c]
    V = tV.cpu().numpy().reshape(n, 3, 3).transpose(0, 2, 1).astype(np.float64)
    S = tS.cpu().numpy().astype(np.float64)
    Fm = F.reshape(n, 3, 3).transpose(0, 2, 1).astype(np.float64)
    nrm = np.maximum(np.abs(Fm).max(axis=(1, 2)), 1.0)
    rec = np.einsum("nij,nj,nkj->nik", U, S, V)
    e = np.abs(rec - Fm).max(axis=(1, 2)) / nrm
    print(f"device svd3 reconstruction: median {np.median(e):.2e}, 99.9 % {np.quantile(e, 0.999):.2e}, worst {e.max():.2e}")
    assert np.median(e) <= 1e-6 and np.quantile(e, 0.999) <= 5e-4 and e.max() <= 5e-3
    eye = np.eye(3)[None]
    assert np.abs(np.einsum("nij,nkj->nik", U, U) - eye).max() <= 5e-6 and np.abs(np.einsum("nij,nkj->nik", V, V) - eye).max() <= 5e-6
    assert np.abs(np.linalg.det(U) - 1).max() <= 1e-5 and np.abs(np.linalg.det(V) - 1).max() <= 1e-5
    assert (np.abs(S[:, 0]) >= np.abs(S[:, 1]) - 1e-6).all() and (np.abs(S[:, 1]) >= np.abs(S[:, 2]) - 1e-6).all()
    exact = np.linalg.svd(Fm, compute_uv=False)
    es = np.abs(np.abs(S) - exact).max(axis=1) / nrm
    er = np.abs(np.abs(S) - np.abs(g["S"].astype(np.float64))).max(axis=1) / nrm
    print(f"singular values: worst vs exact {es.max():.2e}, vs reference {er.max():.2e}")
    assert es.max() <= 1e-4 and er.max() <= 1e-4
    dets = np.linalg.det(Fm)
    assert (np.sign(S[:, 2]) == np.sign(dets))[np.abs(dets) > 1e-4].all()   # the sign lives in the last value


def _exact_fixed_corotated(F32, pb):
    """P F^T vol of the fixed-corotated model in float64 (constitutive_models.cuh:36-73): 2 mu (F - R) F^T + lambda (J - 1) J I with
    R the rotation of the polar decomposition (U, V proper rotations, the sign of det F on the smallest singular value)."""
    F = F32.astype(np.float64).reshape(-1, 3, 3).transpose(0, 2, 1)
    U, S, Vt = np.linalg.svd(F)
    J = np.linalg.det(F)
    U[J < 0, :, 2] *= -1
    R = np.einsum("nij,njk->nik", U, Vt)
    PF = 2 * pb.mu * np.einsum("nij,nkj->nik", F - R, F) * pb.volume + (pb.lambda_ * (J - 1) * J * pb.volume)[:, None, None] * np.eye(3)
    return PF.transpose(0, 2, 1).reshape(-1, 9), J, S


@pytest.mark.parametrize("name,material", [("fc", 1), ("sand", 2), ("nacc", 3)])
def test_device_stress_on_reference_golden(cuda_lib, name, material):
    """compute_stress<M> of the reference (constitutive_models.cuh:36-335, host-compiled, 4000 vectors incl. |F - I| up to ~1 and
    inverted F) against the device functions g2p2g calls.  Tolerance: 3e-5 of the vector's stress scale
    (2 mu + lambda) * volume * max(1, |F|^2) for P F^T, 2e-5 * max(1, |F|) for the returned F, 2e-5 for logJp.
    FIXED_COROTATED: the reference's 4-sweep Jacobi SVD is itself up to 6e-4 of that scale away from the exact value on a few
    vectors (measured against float64), so the device result must be within 3e-5 of the EXACT stress, or no further from the
    reference than 3e-5 + 1.5 x the reference's own error.  SAND / NACC: vectors within rounding of a yield / projection branch point may take the other
    branch: at most 0.5 % may exceed the tolerance (measured: none)."""
    torch = _torch()
    import claymore_b200 as cb
    g = np.load(GOLDEN)
    F = g["F_stress"].astype(np.float32)
    lj = g["log_jp_in"].astype(np.float32)
    n = len(F)
    pb = _default_buffer(cb, material)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tF, tL = d(F), d(lj)
    modes = (0, 1) if material == 1 else (0,)
    fn = np.maximum(np.abs(F).max(axis=1), 1.0)
    scale = (2 * pb.mu + pb.lambda_) * pb.volume * fn ** 2
    if material == 3:
        scale = np.maximum(scale, pb.bm * pb.volume * fn ** 2)
    finite = np.isfinite(g[f"{name}_PF"]).all(axis=1) & np.isfinite(g[f"{name}_F_out"]).all(axis=1)
    for mode in modes:
        oF, oP, oL = torch.zeros(n, 9, device="cuda"), torch.zeros(n, 9, device="cuda"), torch.zeros(n, device="cuda")
        cb._capi.check(cb.lib().cb200_test_stress(material, mode, pb, n, tF.data_ptr(), tL.data_ptr(), oF.data_ptr(), oP.data_ptr(), oL.data_ptr(), None), "test_stress")
        torch.cuda.synchronize()
        Fo, PF, L = oF.cpu().numpy(), oP.cpu().numpy(), oL.cpu().numpy()
        e_pf = np.abs(PF - g[f"{name}_PF"]).max(axis=1) / scale
        e_f = np.abs(Fo - g[f"{name}_F_out"]).max(axis=1) / fn
        if material == 1:
            exact, J, S = _exact_fixed_corotated(F, pb)
            e_dev_exact = np.abs(PF - exact).max(axis=1) / scale
            e_ref_exact = np.abs(g["fc_PF"] - exact).max(axis=1) / scale
            ok = (e_dev_exact <= 3e-5) | (e_pf <= 3e-5 + 1.5 * e_ref_exact)
            proper = J > 1e-3    # away from inversion the polar factor is unique: the device must sit on the exact value
            print(f"fc mode {mode}: device vs exact worst {e_dev_exact[proper].max():.2e} (reference vs exact worst {e_ref_exact[proper].max():.2e}), device vs reference worst {e_pf[finite].max():.2e}; "
                  f"{int((J <= 1e-6).sum())} inverted / singular vectors take the SVD fall-back")
            assert ok[finite].all(), f"fixed-corotated mode {mode}: {int((~ok & finite).sum())} vectors beyond tolerance"
            if mode == 0:
                assert e_dev_exact[proper].max() <= 3e-5
            assert (J <= 1e-6).sum() >= 1, "the golden set must hold inverted F (SVD fall-back of the polar path)"
            assert np.abs(Fo - F).max() == 0.0   # F is not modified by this model
            continue
        bad = finite & ((e_pf > 3e-5) | (e_f > 2e-5) | (np.abs(L - g[f"{name}_log_jp_out"]) > 2e-5))
        print(f"{name}: {finite.sum()} finite vectors, worst PF err {e_pf[finite & ~bad].max():.2e} of scale, worst F err {e_f[finite & ~bad].max():.2e}, outliers {bad.sum()}")
        assert bad.sum() / finite.sum() <= 0.005, f"{name}: {bad.sum()} of {finite.sum()} vectors beyond tolerance"


# ---------------------------------------------------------------------------------------------------------------------------
# capacity: overflow is an error bit, not a memory fault; growth continues the run
# ---------------------------------------------------------------------------------------------------------------------------
def test_block_capacity_overflow_sets_error_bit_only(cuda_lib):
    """A partition that outgrows max_blocks: the step driver must flag kErrBlockCapacity and keep every index inside the
    (max_blocks + 1)-sized containers.  Guard: a second simulator allocated right behind it stays bit-identical to a clean run."""
    # cells [22, 34): 27 particle blocks (block = (cell - 2) >> 2) and 125 exterior blocks; one cell further along -x -y -z the cube
    # straddles 4 blocks per axis: 64 particle blocks, 216 exterior blocks
    scene = scenes.small_cube(lo=22, hi=34, v0=(-2.0, -2.0, -2.0))
    probe = scenes.build_engine(scene, auto_grow=False)
    probe.step(3)
    want = probe.stats()
    href, gref = scenes.grid_by_key(probe.active_keys(), probe.grid())
    probe.close()
    tight = scenes.build_engine(scene, max_blocks=want.exterior_block_count + 2, auto_grow=False)   # fits now, overflows as the cube falls
    guard = scenes.build_engine(scene, auto_grow=False)
    tight.step(200)
    st = tight.stats()
    assert st.error & 1, "the cube has moved several blocks: the partition must have hit the capacity"
    assert st.exterior_block_count <= want.exterior_block_count + 2 and st.neighbor_block_count <= want.exterior_block_count + 2
    guard.step(3)
    hg, gg = scenes.grid_by_key(guard.active_keys(), guard.grid())   # block numbering is atomics-ordered: align by key
    assert np.array_equal(hg, href) and np.abs(gg - gref).max() <= 1e-5 * np.abs(gref).max()
    tight.close()
    guard.close()


def test_growth_from_underprovisioned_start(oracle, cuda_lib):
    """check_capacity's rule (exterior blocks > 3/4 capacity -> x 3/2, gmpm_simulator.cuh:283-300) applied by the step driver itself:
    a run that starts with barely enough blocks grows in place and stays identical to the oracle."""
    scene = scenes.small_cube(v0=(0.5, -3.0, 0.4))
    probe = scenes.build_engine(scene, auto_grow=False)
    ebc0 = probe.stats().exterior_block_count
    probe.close()
    osim = scenes.build_oracle(oracle, scene)
    esim = scenes.build_engine(scene, max_blocks=int(ebc0 * 1.3), auto_grow=True)   # 97 % full at the start: the first poll grows it
    cap0, _ = esim.capacity()
    for k in range(5):
        osim.step(20)
        esim.step(20)
        esim.sync()
    st = esim.stats()
    cap1, events = esim.capacity()
    assert st.error == 0 and events >= 1 and cap1 > cap0, (st.error, events, cap0, cap1)
    assert (st.particle_block_count, st.neighbor_block_count, st.exterior_block_count) == osim.block_counts()
    oh, og = scenes.grid_by_key(osim.active_keys(), osim.grid())
    eh, eg = scenes.grid_by_key(esim.active_keys(), esim.grid())
    assert np.array_equal(oh, eh)
    assert np.abs(eg[:, 0] - og[:, 0]).max() <= 2e-5 * og[:, 0].max() and np.abs(eg[:, 1:] - og[:, 1:]).max() <= 5e-4 * np.abs(og[:, 1:]).max()
    # explicit API: reserve + check_capacity
    esim.reserve(cap1 * 2)
    assert esim.capacity()[0] == cap1 * 2 and esim.check_capacity() == 0
    osim.step(5)
    esim.step(5)
    assert esim.stats().error == 0 and esim.block_counts() == osim.block_counts()
    esim.close()


def test_step_past_frame_horizon_keeps_running(cuda_lib):
    """fps > 0 and plain step(): the frame clock restarts on the device at every frame boundary (the reference's outer frame loop);
    dt must never stay at 0."""
    scene = scenes.small_cube()
    esim = scenes.build_engine(scene, dt=1e-4, fps=240)
    esim.step(130)          # three frames of ~42 sub-steps
    st = esim.stats()
    assert st.error == 0 and st.dt > 0 and st.steps == 130 and 0 <= st.step_time <= 1.0 / 240 + 1e-6
    esim.close()


# ---------------------------------------------------------------------------------------------------------------------------
# MGSP: four ranks, 2 x 2 split, one process (blocks shared by four ranks)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.timeout(600)
def test_mgsp_four_shards_2x2_match_single_domain(oracle, cuda_lib):
    import threading
    from claymore_b200 import mgsp
    scene = scenes.small_cube(v0=(0.3, -1.0, 0.2))
    osim = scenes.build_oracle(oracle, scene)
    world = 4
    sims = []
    for r in range(world):
        part = mgsp.partition_scene_grid(scene, r, world, (2, 2))
        sims.append(mgsp.build_rank_sim(part, r, world, 1e-4, 4000))
    ptrs = [s.mgsp_inbox() for s in sims]
    for s in sims:
        s.mgsp_set_peers(ptrs)
    errs = []

    def run(s):
        try:
            s.initial_setup()
        except Exception as e:  # pragma: no cover
            errs.append(e)
    th = [threading.Thread(target=run, args=(s,)) for s in sims]
    [t.start() for t in th]
    [t.join(180) for t in th]
    assert not errs and not any(t.is_alive() for t in th)

    def check(label):
        oh, og = scenes.grid_by_key(osim.active_keys(), osim.grid())
        lut = {int(h): i for i, h in enumerate(oh)}
        scale = np.abs(og).max(axis=(0, 2), keepdims=True)
        owners = {}
        per_rank = []
        for s in sims:
            st = s.stats()
            assert st.error == 0, (label, st.error)
            k, g = s.active_keys(), s.grid()
            hs = scenes.key_hash(k[: len(g)])
            per_rank.append((hs, g))
            for h in hs:
                owners[int(h)] = owners.get(int(h), 0) + 1
        assert max(owners.values()) == 4, "the 2 x 2 split must produce grid blocks shared by all four ranks"
        seen = np.zeros(len(og), bool)
        for hs, g in per_rank:
            for b, h in enumerate(hs):
                i = lut.get(int(h))
                if i is None:
                    assert np.abs(g[b]).max() == 0, label
                    continue
                if owners[int(h)] > 1:   # every owner holds the full sum
                    assert np.all(np.abs(g[b] - og[i]) <= 2e-4 * scale[0] + 1e-12), (label, f"block shared by {owners[int(h)]} ranks differs from the single-domain block")
                seen[i] = True
        nonzero = np.abs(og).max(axis=(1, 2)) > 0
        assert seen[nonzero].all(), label
        so = osim.particle_state(0)
        se = np.concatenate([s.particle_state(m) for s in sims for m in range(len(s.counts))])
        assert len(so) == len(se)
        idx = scenes.match_particles(so, se, tol=3e-6)
        assert np.abs(se[idx][:, 3:] - so[:, 3:]).max() <= 1e-4, label
        assert len({s.stats().dt for s in sims}) == 1

    check("after setup")
    for k in range(3):
        osim.step(4)
        for s in sims:
            s.step(4)
        for s in sims:
            s.sync()
        check(f"after {4 * (k + 1)} steps")
    for s in sims:
        s.close()


# ---------------------------------------------------------------------------------------------------------------------------
# MGSP: one process per GPU over CUDA IPC (the path SCALE measures); needs a multi-GPU box
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.timeout(900)
@pytest.mark.parametrize("world,split", [(2, "x"), (4, "x"), (4, "2x2")])
def test_mgsp_multiprocess_matches_single_gpu(cuda_lib, world, split):
    torch = _torch()
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, the box has {torch.cuda.device_count()}")
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgsp_worker.py"), "--split", split]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "MGSP_PARITY_OK" in r.stdout, r.stdout[-4000:]


def test_particle_moving_more_than_one_cell_is_dropped_like_the_reference(oracle, cuda_lib):
    """SURVEY.md Appendix B #4 (mgmpm_kernels.cuh:881-885): a particle whose new stencil leaves the 2x2x2 arena of its block keeps
    its state and its bucket entry, but its P2G contribution is lost.  Unreachable with cfl <= 1, so the test raises the CFL factor
    to 2.3 cells per sub-step: about a third of the cube's mass is missing from the grid, identically in the oracle and on the device,
    and the device reports it (error bit 2, kErrLostParticle)."""
    scene = scenes.small_cube(v0=(-3.0, 0.0, 0.0))
    osim = scenes.build_oracle(oracle, scene, dt=1e-1, cfl=2.3)
    esim = scenes.build_engine(scene, dt=1e-1, cfl=2.3, auto_grow=False)
    total = scenes.n_particles(scene) * 1e3 * (1.0 / 64) ** 3 / 8
    for k in range(3):
        osim.step(1)
        esim.step(1)
        st = esim.stats()
        assert (st.particle_block_count, st.neighbor_block_count, st.exterior_block_count) == osim.block_counts()
        oh, og = scenes.grid_by_key(osim.active_keys(), osim.grid())
        eh, eg = scenes.grid_by_key(esim.active_keys(), esim.grid())
        assert np.array_equal(oh, eh)
        mo, me = og[:, 0].sum(dtype=np.float64), eg[:, 0].sum(dtype=np.float64)
        assert mo < 0.8 * total, "the scene must actually drop contributions"
        assert abs(me - mo) <= 1e-5 * mo, (k, me, mo)
        # dt is 20x the elastic time step here: rounding differences grow by ~15x per sub-step (measured 3e-6, 4e-6, 7e-5 of the peak)
        assert np.abs(eg[:, 0] - og[:, 0]).max() <= 3e-4 * og[:, 0].max()
        assert np.abs(eg[:, 1:] - og[:, 1:]).max() <= 2e-3 * np.abs(og[:, 1:]).max()
    assert esim.stats().error & 4
    # nothing is lost from the particle set itself
    assert len(esim.retrieve(0)) == scenes.n_particles(scene)
    esim.close()
