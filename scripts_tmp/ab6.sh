timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
CB200_LIB_VARIANT=t256 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_cube or two_models or jittered or small_max_ppc" > gpurun_out/ab_pytest_t256.log 2>&1; echo "pytest t256 rc=$?"
for v in ne default t256 ne default t256; do
  [ "$v" = "default" ] && vv="" || vv=$v
  CB200_LIB_VARIANT=$vv timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1]); print("VAR[$v]", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["phases_ms"]["g2p2g"], d["phases_ms"]["rebuild"])
except Exception as e: print("VAR[$v] failed", e)
PY
done
