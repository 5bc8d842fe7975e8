#!/bin/bash
# Round-2 evidence on one B200 (run through gpurun from the repo root).  Everything lands in gpurun_out/; the summaries are
# copied into profiles/ by hand.  ncu captures come AFTER the same command has exited 0 without ncu.
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err; echo "bench rc=$?"
python bench.py --workload lat > gpurun_out/r02_bench_lat.json 2> gpurun_out/r02_bench_lat.err; echo "lat rc=$?"
python bench.py --workload c1 --steps 30 --warmup 5 > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err; echo "c1 rc=$?"
python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 rc=$?"
# launch list of the default command (short form of it: one timed step, no CPU leg)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_default_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# the batch search kernel, a launch small enough to replay: 32 queries on 32 slots (8 warps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_team_kernel -c 1 -o gpurun_out/r02_search_team \
    python dev/prof_search.py 1024 32 1 2 8 > gpurun_out/ncu_search.log 2>&1; echo "ncu team rc=$?"
ncu -i gpurun_out/r02_search_team.ncu-rep --page raw --csv > gpurun_out/r02_search_team_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_search_team.ncu-rep --page source --csv > gpurun_out/r02_search_team_source.csv 2>/dev/null
rm -f gpurun_out/r02_search_team.ncu-rep
# DRAM traffic of the batch kernel at a size where the slots are full (metrics only: two passes)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:search_team_kernel -c 1 --csv --log-file gpurun_out/r02_search_team_full_load.csv python dev/prof_search.py 1024 32768 1 2 > gpurun_out/ncu_search_full.log 2>&1; echo "ncu team full rc=$?"
# the field solve kernel on C2 and on C5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:field_solve -s 3 -c 1 -o gpurun_out/r02_field_c2 \
    python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_field.log 2>&1; echo "ncu field c2 rc=$?"
ncu -i gpurun_out/r02_field_c2.ncu-rep --page raw --csv > gpurun_out/r02_field_c2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_field_c2.ncu-rep --page source --csv > gpurun_out/r02_field_c2_source.csv 2>/dev/null
rm -f gpurun_out/r02_field_c2.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:field_solve -s 2 -c 1 --csv --log-file gpurun_out/r02_field_c5.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_field_c5.log 2>&1; echo "ncu field c5 rc=$?"
# traces (instrumented build) with the on-box analysis
for size in 2048 16384; do
  python dev/trace_field.py $size > gpurun_out/r02_trace_$size.log 2>&1
  python dev/analyse_trace.py gpurun_out/trace_$size.bin gpurun_out/trace_${size}_tiles.npz >> gpurun_out/r02_trace_$size.log 2>&1
done
rm -f gpurun_out/*.bin gpurun_out/*_tiles.npz
ls -la gpurun_out
