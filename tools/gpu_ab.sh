#!/bin/bash
# A/B of experiment builds of the library on ONE box (clocks and neighbours differ between boxes, never compare across calls).
# Build variants into claymore_b200/lib/libclaymore_b200_<name>.so (e.g. `nvcc -DCB200_G2P2G_MIN_CTAS=3 ... -o ..._<name>.so`),
# then:  gpurun --timeout 900 -- 'VARIANTS="ctl a default ctl a default" WORKLOADS="spheres5m" bash tools/gpu_ab.sh'
# ("default" = the shipped library).  The full parity suite runs first on the shipped library; every other variant runs a
# parity subset before it is timed (a variant that is fast and wrong must not win an A/B).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
checked=""
for v in ${VARIANTS:-default}; do
  if [ "$v" != "default" ] && [[ " $checked " != *" $v "* ]]; then
    CB200_LIB_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_cube or two_models or jittered or small_max_ppc or differential" > gpurun_out/ab_pytest_$v.log 2>&1
    echo "parity[$v] rc=$? $(tail -1 gpurun_out/ab_pytest_$v.log)"
    checked="$checked $v"
  fi
  for w in ${WORKLOADS:-spheres5m}; do
    [ "$v" = "default" ] && vv="" || vv=$v
    CB200_LIB_VARIANT=$vv timeout 300 python bench.py --workload $w --steps ${STEPS:-200} --warmup 20 --no-cpu-baseline > gpurun_out/ab_${v}_$w.json 2> gpurun_out/ab_${v}_$w.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${v}_$w.json").read().strip().splitlines()[-1])
    print("VAR[$v $w]", round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], "g2p2g", d["phases_ms"]["g2p2g"], "rebuild", d["phases_ms"]["rebuild"])
except Exception as e:
    print("VAR[$v $w] failed", e)
PY
  done
done
