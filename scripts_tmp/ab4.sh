timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/ab_pytest.log
for w in spheres5m spheres40m; do
  timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/ab_p_$w.json 2> gpurun_out/ab_p_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_p_$w.json").read().strip().splitlines()[-1]); print("VAR[$w]", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["gpu_launches"], d["phases_ms"])
except Exception as e: print("VAR[$w] failed", e)
PY
done
