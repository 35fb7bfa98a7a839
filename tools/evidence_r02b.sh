# Round-2 evidence, second part (one B200): final bench record and the surface-QG generator (bench, launch list,
# ncu --set full of the DFT GEMM).
set -x
python bench.py > gpurun_out/bench_c3_g1_r02.json 2> gpurun_out/bench_c3_g1_r02.err
for K in 8 64; do K=$K python tools/sqg_bench.py 2>/dev/null | tail -1 > gpurun_out/sqg_bench_k$K.json; done
N=128 K=32 python tools/sqg_bench.py 2>/dev/null | tail -1 > gpurun_out/sqg_bench_n128_k32.json
N=64 K=64 python tools/sqg_bench.py 2>/dev/null | tail -1 > gpurun_out/sqg_bench_n64_k64.json
K=64 S=2 REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 75 -c 50 --csv --log-file gpurun_out/sqg_launches_r02.csv python tools/sqg_bench.py > gpurun_out/ncu_sqg_l.log 2>&1
K=64 S=1 REPS=1 ncu --set full --clock-control none --import-source on -k regex:sqg_gemm -s 8 -c 4 -o gpurun_out/prof_sqg_gemm_r02 -f python tools/sqg_bench.py > gpurun_out/ncu_sqg.log 2>&1
