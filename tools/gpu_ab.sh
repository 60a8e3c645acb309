#!/bin/bash
# A/B of experiment builds of the library on ONE box (clocks and neighbours differ between boxes, never compare across calls).
# Build variants into claymore_b200/lib/libclaymore_b200_<name>.so (csrc/Makefile `variants`), then e.g.
#   gpurun --timeout 900 -- 'VARIANTS="r01 scalar default r01 scalar default" WORKLOADS="spheres5m spheres40m" bash tools/gpu_ab.sh'
# ("default" = the shipped library).  Every variant runs a parity subset before it is timed (a variant that is fast and wrong
# must not win an A/B); PARITY=0 skips that.
mkdir -p gpurun_out
checked=""
for v in ${VARIANTS:-default}; do
  [ "$v" = "default" ] && vv="" || vv=$v
  if [ "${PARITY:-1}" = "1" ] && [[ " $checked " != *" $v "* ]]; then
    CB200_LIB_VARIANT=$vv timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_cube or two_models or jittered or small_max_ppc or differential or dense" > gpurun_out/ab_pytest_$v.log 2>&1
    echo "parity[$v] rc=$? $(tail -1 gpurun_out/ab_pytest_$v.log)"
    checked="$checked $v"
  fi
  for w in ${WORKLOADS:-spheres5m}; do
    CB200_LIB_VARIANT=$vv timeout 300 python bench.py --workload $w --steps ${STEPS:-200} --warmup 20 --no-cpu-baseline --no-ref-gpu --no-configs1 > gpurun_out/ab_${v}_$w.json 2> gpurun_out/ab_${v}_$w.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${v}_$w.json").read().strip().splitlines()[-1])
    print("VAR[$v $w]", round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], "g2p2g", d["phases_ms"]["g2p2g"], "rebuild", d["phases_ms"]["rebuild"], "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("VAR[$v $w] failed", e)
PY
  done
done
