# Round-1 evidence run (one B200): bench records, ncu launch list, ncu --set full captures.
set -x
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_c3_g1_r01.json 2> gpurun_out/bench_c3_g1_r01.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_ref_r01.json 2>/dev/null
python tools/eig_methods.py > gpurun_out/eig_methods.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
DAB_L96_T=17128 ncu --set full --clock-control none --import-source on -k regex:l96 -s 1 -c 1 -o gpurun_out/prof_forecast_r01 -f python tools/run_forecast_once.py > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"transform_kernel|contract_kernel|stats_reduce" -s 5 -c 4 -o gpurun_out/prof_analysis_r01 -f python tools/run_analysis_once.py > gpurun_out/ncu_an.log 2>&1
python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c5_g1_r01.json 2>/dev/null
python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_g1_r01.json 2>/dev/null
python tools/bench_batched.py > gpurun_out/c2_batched_r01.json 2>/dev/null
