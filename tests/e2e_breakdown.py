"""Where the end-to-end time of bench.py's e2e leg goes (run on a GPU box): create / init_model (H2D) / initial_setup / K steps with a
stats read-back each / retrieve (D2H) / close.   python tests/e2e_breakdown.py [workload] [steps]"""
import sys
import time

sys.path[:0] = ['.', 'tests']
import numpy as np
import torch
import claymore_b200 as cb
from claymore_b200 import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "spheres5m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
scene, _ = scenes.workload(name)
mb = scenes.max_blocks_for(scene)
dx = 1.0 / (1 << scene["domain_bits"])
stream = torch.cuda.Stream()
pinned = [torch.from_numpy(np.ascontiguousarray(m["pos"])).pin_memory() for m in scene["models"]]
outp = [torch.empty_like(p).pin_memory() for p in pinned]
cb.lib()
torch.cuda.synchronize()
for rep in range(2):
    t = [time.perf_counter()]
    cfg = cb.Config(domain_bits=scene["domain_bits"])
    sim = cb.GmpmSimulator(dt=1e-4, fps=0, config=cfg, max_blocks=mb, use_graph=True, stream=stream.cuda_stream)
    t.append(time.perf_counter())
    for m, p in zip(scene["models"], pinned):
        mid = sim.init_model(m["material"], p.numpy(), m["v0"])
        scenes.apply_material(sim, mid, m["material"], dx)
    t.append(time.perf_counter())
    sim.initial_setup()
    t.append(time.perf_counter())
    for _ in range(steps):
        sim.step(1)
        s = sim.stats()
    t.append(time.perf_counter())
    n = sum(len(sim.retrieve(i, out=outp[i].numpy())) for i in range(len(pinned)))
    t.append(time.perf_counter())
    sim.close()
    t.append(time.perf_counter())
    print(name, steps, "steps: create %.1f init_model %.1f setup %.1f steps %.1f retrieve %.1f close %.1f total %.1f ms" % tuple([(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])] + [(t[-1] - t[0]) * 1e3]))
