timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/ab_pytest.log
for v in nm default nm default; do
  [ "$v" = "default" ] && vv="" || vv=$v
  CB200_LIB_VARIANT=$vv timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1]); print("VAR[$v]", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["gpu_launches"], d["phases_ms"]["g2p2g"], d["phases_ms"]["rebuild"])
except Exception as e: print("VAR[$v] failed", e)
PY
done
timeout 200 python bench.py --workload spheres40m --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/ab_40m.json 2> gpurun_out/ab_40m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ab_40m.json").read().strip().splitlines()[-1]); print("VAR[40m]", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["phases_ms"])
PY
