set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1m_launches_bench_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench_f.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:rank_pairs -c 1 -o gpurun_out/prof_rank_pairs_r1m python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_rp_f.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:tc_gemm -c 4 -o gpurun_out/prof_kl_gemms_r1m python tools/one_kl.py 1 > gpurun_out/ncu_kl_f.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:nn_tile -c 1 -o gpurun_out/prof_fast_nn_r1m python tools/bench_fast_nn.py > gpurun_out/ncu_nn_f.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
