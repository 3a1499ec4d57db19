#!/bin/bash
# bench.py at N GPUs under the gradient-exchange variants (see profiles/r02_ddp_overlap.md)
N=${1:-2}
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_ddp_sweep_n${N}_$tag.json 2> gpurun_out/r2_ddp_sweep_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_ddp_sweep_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("$tag: N=%d %.1f samples/s  %.3f ms/step  e2e %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$tag: no result", e)
PY
}
run default
run overlap_sms8 B2POSE_DDP_OVERLAP=1 B2POSE_COMM_SMS=8
run overlap_sms0 B2POSE_DDP_OVERLAP=1 B2POSE_COMM_SMS=0
run fp32_exchange B2POSE_DDP_BF16=0
