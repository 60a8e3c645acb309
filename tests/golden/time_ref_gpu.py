"""Run ON THE GPU BOX: particle-steps/s of the REFERENCE's own kernels on B200 (the "2x" denominator),
same synthetic scenes and material as bench.py.  Wall clock over the reference's host loop (its syncs and D2H counter copies
are part of what a user of the reference waits for; its prints and file output are not issued).  Writes one JSON line."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ref_gpu_binding as rg  # noqa: E402
import scenes  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "spheres5m"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    if which == "sand20m":
        scene = scenes.sand_column()
    elif which == "sand2m":
        scene = scenes.sand_column(domain_bits=8, size=(50, 100, 50))
    else:
        scene = scenes.two_spheres(domain_bits=8 if which == "spheres5m" else 9)
    n = sum(len(m["pos"]) for m in scene["models"])
    sim = rg.build_ref(scene)
    sim.step(5)
    t0 = time.perf_counter()
    ms = sim.time_steps(steps)
    wall = time.perf_counter() - t0
    pbc, nbc, ebc = sim.block_counts()
    print(json.dumps({"impl": "reference claymore kernels, sm_100a build, reference host loop order (no prints / IO)", "workload": which, "particles": n, "steps": steps,
                      "ms_per_step": ms / steps, "wall_ms_per_step": wall * 1e3 / steps, "Mparticle_steps_per_s": n * steps / (ms * 1e-3) / 1e6, "blocks": [pbc, nbc, ebc]}))
    sim.close()


if __name__ == "__main__":
    main()
