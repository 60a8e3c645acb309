"""MGSP host layer: static particle partition across one-process-per-GPU ranks (Projects/MGSP/mgsp.cu:34-96,
mgsp_benchmark.cuh:240-334 in the reference, where one process drives all GPUs with a worker thread each).

torch.distributed is used for plumbing only: rendezvous, the one-time all-gather of the 64-byte CUDA-IPC inbox handles,
barriers and the max-over-ranks of the timings.  The per-sub-step exchange (halo grid blocks, neighbour keys, max |v|^2)
is done by the library's kernels through peer memory; see csrc/mgsp.cuh.
"""
import json
import os
import time

import numpy as np

from . import samplers
from ._capi import Config
from .simulator import GmpmSimulator


def partition_scene(scene, rank, world, axis=0):
    """Static particle partition of the reference's MGSP scenes: every model is cut into `world` equal-count slabs along
    `axis`; rank r owns slab r of every model (one model per device in mgsp.cu:34-81; here one slab of each)."""
    out = []
    for m in scene["models"]:
        parts = samplers.split_slabs(m["pos"], world, axis)
        out.append(dict(material=m["material"], pos=np.ascontiguousarray(parts[rank]), v0=m["v0"]))
    return dict(domain_bits=scene["domain_bits"], models=out)


def partition_scene_global(scene, rank, world, axis=0):
    """Alternative split: all particles of the scene ordered along `axis`, rank r owns the r-th equal-count slab (each
    rank then holds pieces of the models its slab intersects)."""
    allx = np.concatenate([m["pos"][:, axis] for m in scene["models"]])
    cuts = np.quantile(allx, np.linspace(0, 1, world + 1))
    cuts[0], cuts[-1] = -np.inf, np.inf
    out = []
    for m in scene["models"]:
        sel = (m["pos"][:, axis] > cuts[rank]) & (m["pos"][:, axis] <= cuts[rank + 1])
        if sel.any():
            out.append(dict(material=m["material"], pos=np.ascontiguousarray(m["pos"][sel]), v0=m["v0"]))
    return dict(domain_bits=scene["domain_bits"], models=out)


def connect(sim, dist=None):
    """Exchange the inbox IPC handles of all ranks and map them (call on every rank before initial_setup)."""
    if sim.mgsp_world <= 1:
        return
    import torch.distributed as td
    dist = dist or td
    handles = [None] * sim.mgsp_world
    dist.all_gather_object(handles, sim.mgsp_ipc_handle())
    sim.mgsp_open_peers(handles)
    dist.barrier()


def build_rank_sim(scene_part, rank, world, dt, max_blocks, apply_material, stream=None, use_graph=True, max_ppc=128, halo_cap=0):
    cfg = Config(domain_bits=scene_part["domain_bits"], max_ppc=max_ppc)
    sim = GmpmSimulator(dt=dt, fps=0, config=cfg, max_blocks=max_blocks, use_graph=use_graph, stream=stream, mgsp_rank=rank, mgsp_world=world, mgsp_halo_cap=halo_cap)
    dx = 1.0 / (1 << scene_part["domain_bits"])
    for m in scene_part["models"]:
        mid = sim.init_model(m["material"], m["pos"], m["v0"])
        apply_material(sim, mid, m["material"], dx, False)
    return sim


def bench_mgsp(args, scene, label, rank, world, local_rank):
    """N>1 arm of bench.py: weak scaling, max-over-ranks device time, rank 0 prints the JSON line."""
    import torch
    import torch.distributed as dist
    import scenes  # tests/scenes.py: shared material table
    from bench import METRIC, BYTES_BY_MATERIAL, BLOCK_BYTES_G2P2G, ClockSampler, measured_peak_hbm

    part = partition_scene(scene, rank, world)
    n_local = sum(len(m["pos"]) for m in part["models"])
    n_total = sum(len(m["pos"]) for m in scene["models"])
    mb = int(max(4000, n_local / 512 * 5.0))
    stream = torch.cuda.Stream()
    sim = build_rank_sim(part, rank, world, args.dt, mb, scenes.apply_material, stream=stream.cuda_stream, use_graph=not args.no_graph)
    connect(sim, dist)
    sim.initial_setup()
    dist.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        clocks.mark()
    sim.step(args.warmup)
    sim.sync()
    assert sim.stats().error == 0, f"rank {rank}: engine error bits {sim.stats().error} after warm-up"
    l0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    sim.step(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = sim.launch_count - l0
    sim.profile(True)
    sim.step(args.steps)
    phases = sim.profile_phases()
    g_ms, g_n = sim.profile_read()
    sim.profile(False)
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: round(v / args.steps, 4) for k, v in phases.items()})
    clk = clocks.stop() if rank == 0 else None
    st = sim.stats()
    shared, halo_pb = sim.mgsp_halo_counts()
    err = torch.tensor([st.error], device="cuda")
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    assert int(err.item()) == 0, "engine error bits set on some rank"

    # end to end: upload from pinned host memory, setup, K x (step + D2H stats), download -- on every rank, max over ranks
    sim.close()
    pinned = [torch.from_numpy(m["pos"]).pin_memory() for m in part["models"]]
    out_pinned = [torch.empty_like(p).pin_memory() for p in pinned]
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    parts2 = dict(domain_bits=part["domain_bits"], models=[dict(material=m["material"], pos=p.numpy(), v0=m["v0"]) for m, p in zip(part["models"], pinned)])
    sim2 = build_rank_sim(parts2, rank, world, args.dt, mb, scenes.apply_material, stream=stream.cuda_stream, use_graph=not args.no_graph)
    connect(sim2, dist)
    sim2.initial_setup()
    for _ in range(args.steps):
        sim2.step(1)
        s2 = sim2.stats()
    got = sum(len(sim2.retrieve(i, out=out_pinned[i].numpy())) for i in range(len(part["models"])))
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert got == n_local and s2.error == 0
    sim2.close()
    e2e_s = float(t.item())

    if rank == 0:
        material = scene["models"][0]["material"]
        nm = len(part["models"])
        alg = (n_local * BYTES_BY_MATERIAL[material] + st.particle_block_count * BLOCK_BYTES_G2P2G * nm)   # per sub-step on this rank
        peak, kind = measured_peak_hbm()
        per_step_g2p2g_s = g_ms / args.steps * 1e-3
        achieved = alg / per_step_g2p2g_s / 1e9
        out = {
            "metric": METRIC, "value": n_total * args.steps / (ms_total * 1e-3) / 1e6, "unit": "Mparticle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": label, "particles": n_total, "particles_rank0": n_local, "particle_blocks_rank0": st.particle_block_count, "dt": args.dt,
                       "halo_blocks_shared_rank0": shared, "halo_particle_blocks_rank0": halo_pb, "l2": "inputs larger than L2", "graph": not args.no_graph, "phase_ms_per_step_by_rank": gathered,
                       "transport": "kernel stores into CUDA-IPC peer inboxes over NVLink (no NCCL on the data path)"},
            "e2e": {"value": n_total * args.steps / e2e_s / 1e6, "unit": "Mparticle-steps/s", "h2d_bytes_per_step": n_local * 12 / args.steps, "d2h_bytes_per_step": n_local * 12 / args.steps + 76},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "g2p2g_kernel (rank 0, halo + interior launches)", "achieved": achieved, "peak": peak, "peak_kind": kind, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "alg_bytes_per_launch": alg / max(g_n / args.steps, 1), "launches_timed": g_n},
            "cpu_baseline": None,
        }
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
