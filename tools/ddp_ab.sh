#!/bin/bash
# A/B of the overlapped gradient all-reduce at N GPUs: bench.py with B2POSE_DDP_OVERLAP=0 / 1
N=${1:-2}
for ov in 0 1; do
  B2POSE_DDP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + ov)) \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_ddp_n${N}_ov${ov}.json 2> gpurun_out/r2_ddp_n${N}_ov${ov}.err
  echo "overlap=$ov rc=$?"; tail -2 gpurun_out/r2_ddp_n${N}_ov${ov}.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_ddp_n${N}_ov${ov}.json").read().strip().splitlines()[-1])
    print("N=%d overlap=$ov: %.1f samples/s  %.3f ms/step  e2e %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("no result", e)
PY
done
