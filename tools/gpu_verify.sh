#!/bin/bash
# Full single-GPU verification on a B200 box: GPU parity tests, the default bench line (40 M headline scene + reference-GPU leg +
# configs[1] + CPU baseline), one `ncu --set full` capture of the dominant kernel on the headline scene, the per-launch time list,
# and the other single-GPU workloads.  Run from the repo root, e.g.
#   gpurun --timeout 1500 -- 'bash tools/gpu_verify.sh r02'
# Outputs land in gpurun_out/; copy what should be judged into profiles/.
tag=${1:-run}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
timeout 500 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; echo "bench rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:g2p2g -s 6 -c 1 -f -o gpurun_out/prof_g2p2g_40m_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-ref-gpu --no-configs1 > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph --no-ref-gpu --no-configs1 > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launch list rc=$?"
for w in sand20m fluid40m mixed100m; do
  extra=""; [ "$w" = "mixed100m" ] && extra="--no-ref-gpu"
  timeout 500 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-configs1 $extra > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err; echo "$w rc=$?"
done
python - <<PY
import json
for f in ["n1", "sand20m", "fluid40m", "mixed100m"]:
    try:
        d = json.loads(open("gpurun_out/bench_${tag}_%s.json" % f).read().strip().splitlines()[-1])
        rg = d.get("ref_gpu") or {}
        print(f, round(d["value"]), "Mps/s", round(d["ms_per_step"], 4), "ms/step  e2e", round(d["e2e"]["value"]), " roofline", round(d["roofline"]["frac"], 3), d["clocks"]["sm_mhz"], "MHz",
              " ref_gpu", round(rg.get("value", 0)), " x", round(d.get("vs_ref_gpu", 0), 2))
    except Exception as e:
        print(f, "failed:", e)
PY
