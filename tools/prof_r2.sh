#!/bin/bash
# round-2 profile pass: ncu reports stay on the box (too big to pull), CSV digests come back
set -x
O=gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"scan_kernel" -f -o /tmp/r2_prof_cfg python tools/bench_configs.py --only c1,tpch,c5 --iters 2 > $O/r2_prof_cfg.log 2>&1
ncu -i /tmp/r2_prof_cfg.ncu-rep --page raw --csv > $O/r2_prof_cfg_raw.csv 2>/dev/null
for k in 2 4 6 10; do python tools/ncu_lines.py /tmp/r2_prof_cfg.ncu-rep ":::$k" 45 > $O/r2_prof_cfg_lines_$k.txt 2>&1; done
ncu --set full --import-source on --clock-control none -k regex:"pa_" -f -o /tmp/r2_prof_c4 python tools/bench_configs.py --only c4 --iters 2 > $O/r2_prof_c4.log 2>&1
ncu -i /tmp/r2_prof_c4.ncu-rep --page raw --csv > $O/r2_prof_c4_raw.csv 2>/dev/null
for k in 6 8 9 10; do python tools/ncu_lines.py /tmp/r2_prof_c4.ncu-rep ":::$k" 60 > $O/r2_prof_c4_lines_$k.txt 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|finalize_kernel|init_work_table_kernel" -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config > $O/r2_launches.log 2>&1
python bench.py > $O/r2_bench_s3_n1.json 2> $O/r2_bench_s3_n1.err
du -sh $O
