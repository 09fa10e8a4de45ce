#!/bin/bash
# End-of-round evidence (run on the GPU box through gpurun, after the plain commands have exited 0):
#   bash profiles/capture.sh r02
# writes gpurun_out/<tag>_launches_bench.csv (launch list of the bench command) and gpurun_out/<tag>_full.ncu-rep
# (--set full with source of the four hot kernels); the summaries under profiles/ are made from them with
# launch_list.py, ncu_select.py and ncu_lines.py.
set -e
tag=${1:-rXX}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-graphs --skip-cpu > gpurun_out/${tag}_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-graphs --skip-cpu > gpurun_out/${tag}_ncu_launches.log 2>&1
python profiles/run_once.py normal > gpurun_out/${tag}_run_once_plain.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"^target_fused_kernel$" -s 1 -c 1 -f \
    -o gpurun_out/${tag}_full_fused python profiles/run_once.py normal > gpurun_out/${tag}_ncu_full_fused.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"^(segment_kernel|scan_kernel|sample_kernel)$" -s 3 -c 6 -f \
    -o gpurun_out/${tag}_full python profiles/run_once.py normal > gpurun_out/${tag}_ncu_full.log 2>&1
# here: ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_select.py
#       ncu -i X.ncu-rep --page source --csv --kernel-name regex:"^segment_kernel$" > src.csv
#       python profiles/ncu_lines.py src.csv <lib.so> 14segment_kernel   (mangled-name substring: scan_kernelILi11ELb0, target_fused_kernelIx)
