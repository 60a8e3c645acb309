import sys, time; sys.path[:0]=['.','tests']
import numpy as np, torch, scenes, claymore_b200 as cb
scene = scenes.two_spheres(8)
stream = torch.cuda.Stream()
pinned = [torch.from_numpy(np.ascontiguousarray(m["pos"])).pin_memory() for m in scene["models"]]
outp = [torch.empty_like(p).pin_memory() for p in pinned]
cb.lib(); torch.cuda.synchronize()
for rep in range(2):
    t=[time.perf_counter()]
    cfg = cb.Config(domain_bits=8)
    sim = cb.GmpmSimulator(dt=1e-4, fps=0, config=cfg, max_blocks=24437, use_graph=True, stream=stream.cuda_stream); t.append(time.perf_counter())
    for m,p in zip(scene["models"], pinned):
        mid = sim.init_model(m["material"], p.numpy(), m["v0"]); scenes.apply_material(sim, mid, m["material"], 1/256, False)
    t.append(time.perf_counter())
    sim.initial_setup(); t.append(time.perf_counter())
    for _ in range(50): sim.step(1); s=sim.stats()
    t.append(time.perf_counter())
    n=sum(len(sim.retrieve(i, out=outp[i].numpy())) for i in range(2)); t.append(time.perf_counter())
    sim.close(); t.append(time.perf_counter())
    print("create %.1f init_model %.1f setup %.1f steps %.1f retrieve %.1f close %.1f total %.1f ms" % tuple([(b-a)*1e3 for a,b in zip(t[:-1],t[1:])]+[(t[-1]-t[0])*1e3]))
