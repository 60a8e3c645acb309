"""One rank of the multi-process MGSP parity test (launched by tests/test_gpu_scale.py under torch.distributed.run, one process
per GPU): static partition of a two-sphere scene, K sub-steps over CUDA IPC / NVLink, then bench.mgsp_parity -- the union of the
shards against a single-GPU run of the same scene on rank 0 and the agreement of all owners of every shared grid block."""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--split", default="x", choices=["x", "2x2"])
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from claymore_b200 import mgsp, scenes
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # two spheres of 21 cells radius flying at each other, 0.64 M particles; the x-split cuts both, the 2 x 2 split also along y
    scene = scenes.two_spheres(domain_bits=8, radius=0.1645 / 2, centers=((0.40, 0.5, 0.5), (0.60, 0.5, 0.5)), speed=2.0)
    part = mgsp.partition_scene_grid(scene, rank, world, (2, 2)) if a.split == "2x2" else mgsp.partition_scene(scene, rank, world)
    stream = torch.cuda.Stream()
    sim = mgsp.build_rank_sim(part, rank, world, 1e-4, 8000, stream=stream.cuda_stream)
    mgsp.connect(sim, dist)
    sim.initial_setup()
    dist.barrier()
    sim.step(a.steps)
    sim.sync()
    args = types.SimpleNamespace(dt=1e-4, no_single_parity=False)
    parity = bench.mgsp_parity(sim, part, scene, args, rank, world, 8000, 128, stream, a.steps)
    dist.barrier()
    sim.close()
    if rank == 0:
        print("parity:", parity)
        need = 4 if a.split == "2x2" else 2
        assert parity["ok"], parity
        assert parity["key_sets_identical"], parity
        assert parity["max_owners_of_a_block"] >= need, parity
        assert parity["shared_blocks"] > 0
        print("MGSP_PARITY_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
