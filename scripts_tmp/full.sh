timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r01b.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:g2p2g -s 6 -c 1 -f -o gpurun_out/prof_g2p2g_r01_b python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --workload spheres40m --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/b_40m.json 2> gpurun_out/b_40m.err; tail -c 600 gpurun_out/b_40m.json
timeout 300 python bench.py --workload sand20m --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/b_sand20m.json 2> gpurun_out/b_sand.err; tail -c 600 gpurun_out/b_sand20m.json
