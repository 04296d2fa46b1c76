#!/bin/bash
# One GPU call: full-set ncu captures of the three kernels that decide the metric, plus the launch
# list of the bench command (B200_PROFILING.md recipe).  Run on the GPU box (under gpurun, ONE GPU):
#   bash scripts/ncu_capture.sh <tag>        -> gpurun_out/<tag>_*.ncu-rep, gpurun_out/<tag>_launches.csv
# then here:  python scripts/summarize_ncu.py gpurun_out/<tag>_svd_small.ncu-rep > profiles/rNN_svd_small_ncu_full.txt
tag=${1:-cap}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# small-chi SVD: 148 matrices of 128 x 128 (one per SM), second launch
timeout 200 $NCU -k regex:svd_small_kernel -s 1 -c 1 -o gpurun_out/${tag}_svd_small python scripts/prof_svd.py 148 128 2 > gpurun_out/${tag}_svd_small.log 2>&1
# block-Jacobi path at steady state: Gram + tensor-core apply of 50 matrices of 512 x 512
timeout 200 $NCU -k regex:bj_ -s 60 -c 4 -o gpurun_out/${tag}_svd_large python scripts/prof_large.py 50 512 > gpurun_out/${tag}_svd_large.log 2>&1
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
# summaries (the .ncu-rep files stay in gpurun_out/, which is scratch)
for k in svd_small svd_large; do
  python scripts/summarize_ncu.py gpurun_out/${tag}_${k}.ncu-rep > gpurun_out/${tag}_${k}_ncu_full.txt 2>&1
done
python scripts/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt 2>&1
ls -la gpurun_out/${tag}_*
