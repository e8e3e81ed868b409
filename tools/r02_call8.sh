#!/bin/bash
# round 2, GPU call 8 (2 GPUs): both bench arms as the driver launches them at N = 2
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/c8_ref_n2.json 2> $O/c8_ref_n2.err; echo "ref rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > $O/c8_bench_n2.json 2> $O/c8_bench_n2.err; echo "bench rc=$?"; tail -c 800 $O/c8_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c8_bench_n2.json") if l.startswith("{")][-1])
print(d["value"], d["n_gpus"], d["kernel_ms"]); print(json.dumps(d.get("encoder_job")))
PY
