#!/usr/bin/env python
"""bench.py -- particle-steps/s of the fused G2P2G + sparse-grid partition/update sub-step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                  (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                           CPU arm: the oracle port on the host cores

N=1 workload = BASELINE.json configs[1]: GMPM two elastic spheres (fixed-corotated), 256^3 grid, ~5 M particles.
N>1: MGSP static particle partition; weak scaling -- each rank owns ~5 M particles of a two-sphere scene of N x 5 M
particles (radius scaled by N^(1/3); 512^3 grid for N >= 4), halo grid blocks exchanged over NCCL every sub-step.
A "step" is one sub-step (grid update, g2p2g, partition rebuild) with fixed dt (default_dt 1e-4; the CFL bound never binds).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "million particle-steps/sec"
FC_BYTES_G2P2G = 104.0          # algorithmic bytes per particle-step, g2p2g kernel, fixed-corotated
BLOCK_BYTES_G2P2G = 2048.0      # + per particle block: 768 B velocity read + 1024 B accumulate-write + 256 B cell counters
BYTES_BY_MATERIAL = {0: 40.0, 1: 104.0, 2: 112.0, 3: 112.0}


def workload(name, n_gpus):
    import scenes
    if name == "spheres5m":
        if n_gpus == 1:
            return scenes.two_spheres(domain_bits=8), "GMPM two elastic spheres (fixed-corotated), 256^3 grid, 5M particles"
        bits = 8 if n_gpus <= 2 else 9
        # same number of particles per rank: radius in cells scales with N^(1/3); keep the cell size of the chosen grid
        r_cells = 42.1 * n_gpus ** (1.0 / 3.0)
        dx = 1.0 / (1 << bits)
        r = r_cells * dx
        gap = 0.05
        c0, c1 = 0.5 - gap / 2 - r, 0.5 + gap / 2 + r
        sc = scenes.two_spheres(domain_bits=bits, radius=r, centers=((c0, 0.5, 0.5), (c1, 0.5, 0.5)))
        return sc, f"MGSP two elastic spheres (fixed-corotated), {1 << bits}^3 grid, {n_gpus}x5M particles, x-slab static partition"
    if name == "spheres40m":
        return scenes.two_spheres(domain_bits=9), "GMPM two elastic spheres (fixed-corotated), 512^3 grid, 40M particles"
    if name == "sand20m":
        return scenes.sand_column(), "GMPM sand column collapse (Drucker-Prager), 512^3 grid, 20M particles"
    if name == "sand2m":
        return scenes.sand_column(domain_bits=8, size=(50, 100, 50)), "sand column (Drucker-Prager), 256^3 grid, 2M particles"
    if name == "fluid40m":
        return scenes.fluid_dam(), "weakly-compressible fluid dam break, 1024^3 grid, 40M particles"
    if name == "fluid5m":
        return scenes.fluid_dam(domain_bits=9, size=(100, 62, 100)), "weakly-compressible fluid dam, 512^3 grid, 5M particles"
    if name == "cube140k":
        return scenes.jelly_cube(), "jelly cube (fixed-corotated), 128^3 grid, 140608 particles"
    if name == "spheres640k":
        return scenes.two_spheres(domain_bits=8, radius=0.1645 / 2), "two elastic spheres (fixed-corotated), 256^3 grid, r=21 cells, 0.64M particles"
    raise SystemExit(f"unknown workload {name}")


def max_blocks_for(scene):
    n = sum(len(m["pos"]) for m in scene["models"])
    return int(max(4000, n / 512 * 2.5))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (recipe in /opt/skills/guides/B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        t0 = time.time()
        while not self.rows and time.time() - t0 < 3.0:   # the first sample takes a few hundred ms: wait for it
            time.sleep(0.01)
        self.first = len(self.rows)

    def mark(self):
        """Samples from here on are 'under load' (warm-up, timed region, per-kernel timing pass)."""
        self.first = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows[max(getattr(self, "first", 0) - 1, 0):]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_threads():
    """Threads for the CPU legs.  The oracle port stops scaling past 16 threads (measured on the 128-core B200 host with the
    0.63 M-particle sample: 8 -> 8.6, 16 -> 16.1, 32 -> 14.4, 64 -> 11.2, 128 -> 1.7 M particle-steps/s), so 16 is "all it can use"."""
    return max(1, min(os.cpu_count() or 1, 16))


def time_cpu_port(scene, seconds_budget, threads, max_steps=None):
    """The oracle port on the host cores: (particle-steps/s in millions, particles, steps)."""
    import oracle_binding as ob
    import scenes
    n = sum(len(m["pos"]) for m in scene["models"])
    osim = scenes.build_oracle(ob, scene, max_blocks=max_blocks_for(scene), threads=threads)
    osim.step(1)
    t0 = time.perf_counter()
    steps = 0
    while True:
        osim.step(1)
        steps += 1
        el = time.perf_counter() - t0
        if el >= seconds_budget or (max_steps and steps >= max_steps):
            break
    osim.close()
    return n * steps / el / 1e6, n, steps


def run_reference(args):
    """--impl reference: the reference has no CPU implementation of this path (Projects/TaichiScripts/gmpm.py is a
    57-line stub, SURVEY.md 'Read this first' #4), so the arm times the oracle port (OpenMP) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_binding as ob
    import scenes
    cores = cpu_threads()
    scene, label = workload("spheres640k", 1)
    n = sum(len(m["pos"]) for m in scene["models"])
    osim = scenes.build_oracle(ob, scene, max_blocks=max_blocks_for(scene), threads=cores)
    for _ in range(max(args.warmup, 1)):
        osim.step(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        osim.step(1)
    el = time.perf_counter() - t0
    osim.close()
    v = n * args.steps / el / 1e6
    _, wl_label = workload(args.workload, args.gpus)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mparticle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_label, "sample": label},
        "cpu_baseline": {"value": v, "unit": "Mparticle-steps/s", "cores": cores, "kind": "port", "sample": f"{label}: {n} particles x {args.steps} sub-steps per timed run (1/8 of the particles of the N=1 workload, same grid/material/dt)"},
        "e2e": {"value": v, "unit": "Mparticle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_b200(args):
    # rank 0 prints exactly one JSON line on stdout: NCCL's banner / debug lines (the box exports NCCL_DEBUG) go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # this level alone printf()s a banner to stdout
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    import scenes
    import claymore_b200 as cb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cb.lib()  # fail loudly if the CUDA library is missing

    scene, label = workload(args.workload, world)
    if world > 1:
        from claymore_b200 import mgsp
        return mgsp.bench_mgsp(args, scene, label, rank, world, local_rank)

    n_particles = sum(len(m["pos"]) for m in scene["models"])
    material = scene["models"][0]["material"]
    dx = 1.0 / (1 << scene["domain_bits"])
    mb = max_blocks_for(scene)

    stream = torch.cuda.Stream()  # the kernels are launched on this stream; the timing events are recorded on it too

    def fresh(use_graph):
        return scenes.build_engine(scene, dt=args.dt, max_blocks=mb, use_graph=use_graph, stream=stream.cuda_stream)

    # ---- device-resident timing: K sub-steps, CUDA events, graph replay --------------------------------------
    sim = fresh(use_graph=not args.no_graph)
    clocks = ClockSampler(local_rank)
    clocks.start()
    clocks.mark()
    sim.step(args.warmup)
    sim.sync()
    st0 = sim.stats()
    assert st0.error == 0, f"engine error bits {st0.error} after warm-up"
    l0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    sim.step(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = sim.launch_count - l0
    # ---- per-kernel timing for the roofline: same simulation, next K sub-steps, event pair around every g2p2g launch
    sim.profile(True)
    sim.step(args.steps)
    g2p2g_ms, g2p2g_launches = sim.profile_read()
    phases_ms = {k: round(v / args.steps, 4) for k, v in sim.profile_phases().items()}  # ungraphed pass, event per phase
    sim.profile(False)
    clk = clocks.stop()
    st = sim.stats()
    assert st.error == 0, f"engine error bits {st.error}"
    pbc = st.particle_block_count
    sim.close()

    value = n_particles * args.steps / (ms_total * 1e-3) / 1e6
    per_model = [len(m["pos"]) for m in scene["models"]]
    # algorithmic bytes of one sub-step's g2p2g work = particles x B_p + 2048 B per particle block that holds particles of a
    # model; models of one material are handled by ONE launch, so launches per step = distinct materials
    launches_per_step = max(g2p2g_launches, 1) / args.steps
    alg_bytes_per_launch = (sum(per_model) * BYTES_BY_MATERIAL[material] + pbc * BLOCK_BYTES_G2P2G) / launches_per_step
    avg_launch_s = g2p2g_ms / max(g2p2g_launches, 1) * 1e-3
    peak, peak_kind = measured_peak_hbm()
    achieved = alg_bytes_per_launch / avg_launch_s / 1e9
    roofline = {"bound": "hbm", "kernel": "g2p2g_kernel<%s>" % {0: "J_FLUID", 1: "FIXED_COROTATED", 2: "SAND", 3: "NACC"}[material], "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "alg_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_launch_s * 1e3,
                "launches_timed": g2p2g_launches, "share_of_step": (g2p2g_ms / args.steps) / (ms_total / args.steps)}
    traffic_file = os.path.join(ROOT, "profiles", "g2p2g_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- end to end through the public API with HOST buffers inside the timed region ---------------------------
    # upload of every model from pinned host memory (init_model), initial_setup, K sub-steps each followed by a
    # device->host read of the step result (block counts, dt, max velocity), and the per-frame particle download.
    pinned = [torch.from_numpy(np.ascontiguousarray(m["pos"])).pin_memory() for m in scene["models"]]
    out_pinned = [torch.empty_like(p).pin_memory() for p in pinned]   # the caller's output buffers, reused frame after frame
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cfg = cb.Config(domain_bits=scene["domain_bits"])
    sim2 = cb.GmpmSimulator(dt=args.dt, fps=0, config=cfg, max_blocks=mb, use_graph=not args.no_graph, stream=stream.cuda_stream)
    for m, p in zip(scene["models"], pinned):
        mid = sim2.init_model(m["material"], p.numpy(), m["v0"])
        scenes.apply_material(sim2, mid, m["material"], dx, False)
    sim2.initial_setup()
    stats_bytes = 0
    for _ in range(args.steps):
        sim2.step(1)
        s = sim2.stats()
        stats_bytes += 76
    out_n = 0
    for i in range(len(scene["models"])):
        out_n += len(sim2.retrieve(i, out=out_pinned[i].numpy()))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert out_n == n_particles and s.error == 0
    sim2.close()
    e2e_value = n_particles * args.steps / e2e_s / 1e6
    e2e = {"value": e2e_value, "unit": "Mparticle-steps/s", "h2d_bytes_per_step": n_particles * 12 / args.steps, "d2h_bytes_per_step": n_particles * 12 / args.steps + stats_bytes / args.steps,
           "note": "timed: init_model H2D from pinned host, initial_setup, K x (step + D2H stats), retrieve D2H of all positions"}

    # ---- CPU baseline beside it: the oracle port on the host cores, bounded sample -------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        cores = cpu_threads()
        sc, lab = workload("spheres640k", 1)
        v, n_s, steps_s = time_cpu_port(sc, seconds_budget=args.cpu_seconds, threads=cores)
        cpu = {"value": v, "unit": "Mparticle-steps/s", "cores": cores, "kind": "port", "sample": f"{lab}: {n_s} particles x {steps_s} sub-steps (~{args.cpu_seconds:.0f} s), oracle port (OpenMP)"}

    out = {
        "metric": METRIC, "value": value, "unit": "Mparticle-steps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label, "particles": n_particles, "particle_blocks": pbc, "dt": args.dt, "l2": "inputs larger than L2 (particle bins >= 500 MB)",
                   "graph": not args.no_graph},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
        "phases_ms": phases_ms,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="spheres5m")
    ap.add_argument("--dt", type=float, default=1e-4)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
