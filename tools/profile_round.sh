#!/bin/bash
# Round evidence captured on a GPU box (through gpurun); summaries are copied into profiles/ by hand afterwards.
#   tools/profile_round.sh <tag>       e.g. r2a
set -x
T=${1:-r2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_bench_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --quick --no-graph > gpurun_out/ncu_bench_${T}.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:rank_pairs -c 1 \
    -o gpurun_out/prof_rank_pairs_${T} python tools/time_rank.py --iters 1 > gpurun_out/ncu_rp_${T}.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:tc_gemm|kl_dz_fast|kl_prep_fast" -s 20 -c 5 \
    -o gpurun_out/prof_kl_${T} python tools/time_kl.py --iters 1 > gpurun_out/ncu_kl_${T}.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
