#!/bin/bash
# Round-end evidence on an 8-GPU box: replica tests, bench at 8 and 4 GPUs, configs[1] through the CLI on 8 GPUs, host packing rate.
set -x
R=${1:-r02}
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 \
    > gpurun_out/${R}_bench_n$n.out 2> gpurun_out/${R}_bench_n$n.err
  grep '^{' gpurun_out/${R}_bench_n$n.out > gpurun_out/${R}_bench_n$n.json
done
g++ -O3 -mavx2 -pthread -o /tmp/pack_bench tools/pack_bench.cc && /tmp/pack_bench 300 1 8 16 32 64 > gpurun_out/${R}_pack_bench_8gpu_host.jsonl
nproc
python tools/file_pipeline_bench.py --pairs 10000000 --threads $(nproc) --gpus 8 --variants gz-gz,plain-plain,bgz-plain > gpurun_out/${R}_file_pipeline_n8.json 2> gpurun_out/${R}_file_pipeline_n8.err
tail -c 1500 gpurun_out/${R}_file_pipeline_n8.err
