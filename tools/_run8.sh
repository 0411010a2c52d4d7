R=r02
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_files.py -m gpu -q 2>&1 | tail -3
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 \
    > gpurun_out/${R}_bench_n$n.out 2> gpurun_out/${R}_bench_n$n.err
  grep '^{' gpurun_out/${R}_bench_n$n.out > gpurun_out/${R}_bench_n$n.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_bench_n$n.json").read())
print($n, d["value"], d["e2e"]["value"], d["parity_vs_oracle"], [(w["workload"][17:27], w["gbp_s"]) for w in d["workloads"]])
PY
done
nproc
