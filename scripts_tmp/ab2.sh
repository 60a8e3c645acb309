timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
CB200_LIB_VARIANT=r3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_cube or two_models or jittered or small_max_ppc or differential" > gpurun_out/ab_pytest_r3.log 2>&1; echo "pytest r3 rc=$?"
for v in $VARIANTS; do
  [ "$v" = "default" ] && vv="" || vv=$v
  CB200_LIB_VARIANT=$vv timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1]); print("VAR[$v]", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
except Exception as e: print("VAR[$v] failed", e)
PY
done
