#!/bin/bash
# usage (on the GPU box): tools/gpu_ab.sh OUTTAG variant1 variant2 ...   -> gpurun_out/ab_OUTTAG.jsonl
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  NH_LIB_PATH=$PWD/nohuman_b200/variants/libnh_$v.so timeout 600 python tools/kernel_ab.py --tag $v $AB_ARGS >> gpurun_out/ab_$tag.jsonl 2>> gpurun_out/ab_$tag.err
done
python - <<PY
import json
for ln in open("gpurun_out/ab_$tag.jsonl"):
    d=json.loads(ln)
    print(d["tag"], " ".join(f"{k}:{v['ms_fused']:.3f}/{v['gbp_s']:.0f}/{v['sha'][:6]}" for k,v in d["cases"].items()))
PY
