# Round-2 evidence run (one B200): bench record, ncu launch list, ncu --set full captures of every kernel of the cycle.
set -x
python bench.py > gpurun_out/bench_c3_g1_r02.json 2> gpurun_out/bench_c3_g1_r02.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_ref_r02.json 2>/dev/null
# launch list: -s skips the set-up launches (workload generation, the forecast tuner's 60 launches)
DAB_L96_T=16128 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 240 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches_r02.log 2>&1
# K1: the first design at C3 (the tuner's pick there), the pipelined chains at C4 (its pick there)
DAB_L96_T=16128 ncu --set full --clock-control none --import-source on -k regex:l96_rk4 -s 1 -c 1 -o gpurun_out/prof_forecast_c3_r02 -f python tools/run_forecast_once.py > gpurun_out/ncu_f_c3.log 2>&1
N=4194304 DAB_L96_V6=1 ncu --set full --clock-control none --import-source on -k regex:l96_v6 -s 1 -c 1 -o gpurun_out/prof_forecast_c4_r02 -f python tools/run_cycler_once.py > gpurun_out/ncu_f_c4.log 2>&1
# analysis kernels at C3
ncu --set full --clock-control none --import-source on -k regex:"transform_kernel|contract_kernel|stats_reduce|ns_mc_transform" -s 4 -c 4 -o gpurun_out/prof_analysis_c3_r02 -f python tools/run_analysis_once.py > gpurun_out/ncu_an_c3.log 2>&1
