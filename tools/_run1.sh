timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_filter_default.log 2>&1; echo "default rc=$?"; tail -2 gpurun_out/pytest_filter_default.log | cut -c1-400
NH_FILTER_MODE=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_db_variants.py tests/test_gpu_synth.py -m gpu -x -q > gpurun_out/pytest_filter_mode2.log 2>&1; echo "mode2 rc=$?"; tail -2 gpurun_out/pytest_filter_mode2.log | cut -c1-400
for m in 0 1; do
NH_FILTER_MODE=$m ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_requests_srcunit_tex.sum,dram__bytes_read.sum \
  --clock-control none --profile-from-start off -k regex:k_stream -c 2 --csv --log-file gpurun_out/ncu_filter_$m.csv python bench.py --steps 1 --warmup 1 --launches-per-step 1 --no-e2e --no-cpu-baseline --no-workloads --no-parity > gpurun_out/ncu_filter_$m.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/ncu_filter_$m.csv") if not l.startswith("=="))]
h=rows[0]; mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
last=max(r[ii] for r in rows[1:])
print("mode $m", {r[mi].split('.')[0][-28:]: r[vi] for r in rows[1:] if r[ii]==last})
PY
done
for m in 0 1 2; do
  NH_FILTER_MODE=$m timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_filter_$m.json 2> gpurun_out/bench_filter_$m.err
  python - <<PY
import json
for l in open("gpurun_out/bench_filter_$m.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("mode $m value", d["value"], "ms/launch", d["roofline"]["ms_per_launch"], "parity", d["parity_vs_oracle"]["mismatches"], "req/lookup", d["roofline_probe"]["sectors_per_lookup"])
        print("   ", [(w["workload"][17:27], w["gbp_s"], w["sectors_per_lookup"], w["parity_vs_oracle"]["mismatches"] if "parity_vs_oracle" in w else None) for w in d["workloads"]])
PY
done
