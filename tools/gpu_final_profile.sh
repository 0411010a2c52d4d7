#!/bin/bash
# Round-end evidence on one B200: bench line, reference arm, ncu launch list, ncu --set full with sources,
# ncu of the roofline helper.  Everything lands in gpurun_out/ (copied into profiles/ by hand).
set -x
R=${1:-r02}
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/${R}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-workloads --no-parity > gpurun_out/${R}_launches.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"k_stream|k_score|k_plan" -c 6 -f \
    -o gpurun_out/${R}_full python bench.py --steps 1 --warmup 1 --launches-per-step 1 --no-e2e --no-cpu-baseline --no-workloads --no-parity > gpurun_out/${R}_full.log 2>&1
ncu --metrics gpu__time_duration.sum,lts__t_requests_srcunit_tex.sum,lts__t_sectors_srcunit_tex.sum,dram__bytes_read.sum \
    --clock-control none -k regex:k_probe_pattern --csv --log-file gpurun_out/${R}_probe_pattern_ncu.csv python tools/probe_pattern_ncu.py > gpurun_out/${R}_probe_pattern.json 2> gpurun_out/${R}_probe_pattern.err
ls -la gpurun_out/${R}_*
