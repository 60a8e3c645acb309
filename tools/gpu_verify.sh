#!/bin/bash
# Full verification on a B200 box: GPU parity tests, the default bench line, one `ncu --set full` capture of the dominant
# kernel, the per-launch time list, and the other single-GPU workloads.  Run from the repo root, e.g.
#   gpurun --timeout 1500 -- 'bash tools/gpu_verify.sh r02'
# Outputs land in gpurun_out/; copy what should be judged into profiles/.
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
timeout 300 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:g2p2g -s 6 -c 1 -f -o gpurun_out/prof_g2p2g_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launch list rc=$?"
for w in spheres40m sand20m fluid40m; do
  timeout 300 python bench.py --workload $w --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_$w.json 2> /dev/null
done
python - <<PY
import json
for f in ["n1", "spheres40m", "sand20m", "fluid40m"]:
    try:
        d = json.loads(open("gpurun_out/bench_${tag}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "Mps/s", round(d["ms_per_step"], 4), "ms/step  e2e", round(d["e2e"]["value"]), " roofline", round(d["roofline"]["frac"], 3), d["clocks"]["sm_mhz"], "MHz")
    except Exception as e:
        print(f, "failed:", e)
PY
