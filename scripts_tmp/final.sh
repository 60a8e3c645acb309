timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 300 python bench.py > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err; echo "bench rc=$?"; wc -l gpurun_out/bench_final2_n1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:g2p2g -s 6 -c 1 -f -o gpurun_out/prof_g2p2g_r01_final2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01_final2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu2 rc=$?"
timeout 300 python bench.py --workload sand20m --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final2_sand20m.json 2> /dev/null
timeout 300 python bench.py --workload fluid40m --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final2_fluid40m.json 2> /dev/null
python -c "
import json
for f in ['n1','sand20m','fluid40m']:
    d=json.loads(open('gpurun_out/bench_final2_%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['roofline']['frac'], d['clocks']['sm_mhz'])
"
