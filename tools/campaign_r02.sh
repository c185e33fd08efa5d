# Evidence campaign of round 2 (one GPU): what produced the files under profiles/r02_*.  Run through gpurun from the repo root.
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/f_gputest.log 2>&1; tail -2 gpurun_out/f_gputest.log
python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo bench rc $?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo ref rc $?
python tools/bench_configs.py > gpurun_out/f_configs.jsonl 2> gpurun_out/f_configs.err
python tools/resident_sweep.py --cases 20000x6,20000x12,50000x24,31250x96,62500x96,93750x96,125000x96,250000x96,500000x96,1000000x96 --steps 1000 > gpurun_out/f_resident_sizes.jsonl 2>&1
( echo "# python tools/ps_rows_sweep.py 200000 48 6  (no control genes)"; timeout 200 python tools/ps_rows_sweep.py 200000 48 6; echo "# python tools/ps_rows_sweep.py 200000 48 6 0 controls  (half of the genes on a second design, cfg4)"; timeout 200 python tools/ps_rows_sweep.py 200000 48 6 0 controls; echo "# python tools/ps_rows_sweep.py 200000 48 6 0.02  (2 % missing values: most rows take the general path)"; timeout 200 python tools/ps_rows_sweep.py 200000 48 6 0.02 ) > gpurun_out/f_ps_rows_sweeps.log 2>&1
FN_BENCH_SKIP_CPU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/f_bench_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/f_ncu_bench.log 2>&1; echo ncu-bench rc $?
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -o gpurun_out/f_scale1 -f python tools/profile_eval.py 1000000 96 Model_t 4 > gpurun_out/f_ncu_scale1.log 2>&1; echo rc $?
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -o gpurun_out/f_ps_rows -f python tools/profile_eval.py 200000 48 Model_t_per_sample 4 6 > gpurun_out/f_ncu_ps.log 2>&1; echo rc $?
ncu --set full --clock-control none --import-source on -k regex:prepare_kernel -c 1 -o gpurun_out/f_prepare -f python tools/profile_eval.py 1000000 96 Model_t 2 > gpurun_out/f_ncu_prep.log 2>&1; echo rc $?
