timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
for v in $VARIANTS; do
  for w in spheres5m sand20m; do
  [ "$v" = "default" ] && vv="" || vv=$v
  CB200_LIB_VARIANT=$vv timeout 200 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/ab_${v}_$w.json 2> gpurun_out/ab_${v}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${v}_$w.json").read().strip().splitlines()[-1]); print("VAR[$v $w]", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["phases_ms"])
except Exception as e: print("VAR[$v $w] failed", e)
PY
  done
done
