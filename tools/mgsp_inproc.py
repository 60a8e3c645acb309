#!/usr/bin/env python
"""Debug aid: N MGSP ranks inside ONE process on ONE GPU (peers wired with set_peers instead of CUDA IPC), global x-slab split of a
named workload, K sub-steps, then the union of the shards against a single-domain engine run.  Not under compute-sanitizer: it
serialises kernel launches, and the ranks' flag waits need their peers' kernels to run concurrently (the run hangs).
    python tools/mgsp_inproc.py --workload spheres640k --ranks 4 --steps 20 [--split global|model|2x2]"""
import argparse
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="spheres640k")
    ap.add_argument("--ranks", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--split", default="global")
    ap.add_argument("--max-ppc", type=int, default=128)
    ap.add_argument("--check-every", type=int, default=0)
    ap.add_argument("--diagnose", action="store_true")
    a = ap.parse_args()
    from claymore_b200 import mgsp, scenes
    scene, label = scenes.workload(a.workload)
    world = a.ranks
    sims, parts = [], []
    for r in range(world):
        if a.split == "global":
            parts.append(mgsp.partition_scene_global(scene, r, world))
        elif a.split == "2x2":
            parts.append(mgsp.partition_scene_grid(scene, r, world, (2, 2)))
        else:
            parts.append(mgsp.partition_scene(scene, r, world))
    mb = int(max(4000, max(scenes.n_particles(p) for p in parts) / 512 * 5.0))   # the SAME capacity on every rank
    for r, part in enumerate(parts):
        print(f"rank {r}: {scenes.n_particles(part)} particles, {len(part['models'])} models, max_blocks {mb}", flush=True)
        sims.append(mgsp.build_rank_sim(part, r, world, 1e-4, mb, max_ppc=a.max_ppc))
    ptrs = [s.mgsp_inbox() for s in sims]
    for s in sims:
        s.mgsp_set_peers(ptrs)
    errs = []

    def run(s):
        try:
            s.initial_setup()
        except Exception as e:
            errs.append(e)
    th = [threading.Thread(target=run, args=(s,)) for s in sims]
    [t.start() for t in th]
    [t.join(300) for t in th]
    assert not errs and not any(t.is_alive() for t in th), errs
    print("setup ok", flush=True)
    def owners(label):
        allh, allg, rk, idx, pbcs, allkeys = [], [], [], [], [], []
        for r, s in enumerate(sims):
            s.sync()
            st = s.stats()
            k, g = s.active_keys()[: st.neighbor_block_count], s.grid()
            allh.append(scenes.key_hash(k))
            allg.append(g)
            rk.append(np.full(len(k), r))
            idx.append(np.arange(len(k)))
            pbcs.append(st.particle_block_count)
            allkeys.append(k)
        h = np.concatenate(allh)
        g = np.concatenate(allg, 0)
        if a.diagnose:
            rk_, idx_, keys_ = np.concatenate(rk), np.concatenate(idx), np.concatenate(allkeys, 0)
            o = np.lexsort((-np.abs(g).sum(axis=(1, 2)), h))
            hs, gs = h[o], g[o]
            first = np.ones(len(hs), bool)
            first[1:] = hs[1:] != hs[:-1]
            fresh = ~first & (np.abs(gs).max(axis=(1, 2)) == 0)
            ii = np.nonzero(fresh)[0][:12]
            for i in ii:
                j = i - 1   # the fullest copy of the same key
                while not first[j]:
                    j -= 1
                print(f"   key {keys_[o][i]} zero on rank {rk_[o][i]} (block {idx_[o][i]}, pbc {pbcs[rk_[o][i]]}); full on rank {rk_[o][j]} (block {idx_[o][j]}, pbc {pbcs[rk_[o][j]]}) mass {gs[j, 0].sum():.3e} max cell {gs[j, 0].max():.3e}")
        o = np.lexsort((-np.abs(g).sum(axis=(1, 2)), h))   # per key the fullest copy first
        hs, gs = h[o], g[o]
        first = np.ones(len(hs), bool)
        first[1:] = hs[1:] != hs[:-1]
        grp = np.cumsum(first) - 1
        fresh = ~first & (np.abs(gs).max(axis=(1, 2)) == 0) & (np.abs(gs[first][grp]).max(axis=(1, 2)) > 0)   # new on that rank in this sub-step: not yet tagged, never read (see bench.mgsp_parity)
        bad = (np.abs(gs - gs[first][grp]).max(axis=(1, 2)) > 1e-4 * np.abs(gs).max()) & ~fresh
        print(label, "shared", int((~first).sum()), "fresh", int(fresh.sum()), "disagreeing", int(bad.sum()), "blocks", [s.block_counts() for s in sims], flush=True)
        return int(bad.sum())

    for k in range(a.steps):
        for s in sims:
            s.step(1)
        if a.check_every and (k + 1) % a.check_every == 0:
            if owners(f"after step {k + 1}:"):
                break
    for r, s in enumerate(sims):
        s.sync()
        st = s.stats()
        print(f"rank {r}: blocks {st.particle_block_count}/{st.neighbor_block_count}/{st.exterior_block_count} error {st.error} halo {s.mgsp_halo_counts()}", flush=True)
        assert st.error == 0
    single = scenes.build_engine(scene, max_blocks=scenes.max_blocks_for(scene), max_ppc=a.max_ppc, auto_grow=False)
    single.step(a.steps)
    sst = single.stats()
    sh, sg = scenes.grid_by_key(single.active_keys()[: sst.neighbor_block_count], single.grid())
    allh, allg = [], []
    for s in sims:
        st = s.stats()
        k, g = s.active_keys()[: st.neighbor_block_count], s.grid()
        allh.append(scenes.key_hash(k))
        allg.append(g)
    h = np.concatenate(allh)
    g = np.concatenate(allg, 0)
    o = np.lexsort((-np.abs(g).sum(axis=(1, 2)), h))
    hs, gs = h[o], g[o]
    first = np.ones(len(hs), bool)
    first[1:] = hs[1:] != hs[:-1]
    grp = np.cumsum(first) - 1
    fresh = ~first & (np.abs(gs).max(axis=(1, 2)) == 0) & (np.abs(gs[first][grp]).max(axis=(1, 2)) > 0)
    owner_err = (np.abs(gs - gs[first][grp]).max(axis=(1, 2))[~fresh]).max() / np.abs(gs).max()
    print("shared blocks", int((~first).sum()), "max owners", int(np.bincount(grp).max()), "owner err", owner_err)
    assert np.array_equal(hs[first], sh), "key sets differ"
    scale = np.array([np.abs(sg[:, 0]).max()] + [np.abs(sg[:, 1:]).max()] * 3)
    err = np.abs(gs[first] - sg).max(axis=(0, 2)) / scale
    print("union vs single, per channel max err / scale:", err)
    bad = (np.abs(gs - gs[first][grp]).max(axis=(1, 2)) > 1e-4 * np.abs(gs).max()) & ~fresh
    print("blocks whose owners disagree:", int(bad.sum()), "of shared", int((~first).sum()), "fresh", int(fresh.sum()))
    assert owner_err <= 1e-4 and (err <= 1e-3).all()
    print("MGSP_INPROC_OK")


if __name__ == "__main__":
    main()
