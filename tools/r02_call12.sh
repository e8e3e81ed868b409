#!/bin/bash
# round 2, GPU call 12 (8 GPUs): BASELINE config 4 (3840x2160 x 600 frames, 8 GOPs) sharded over 8 / 4 / 2 GPUs of one box
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
nproc > $O/c12_box.txt; nvidia-smi -L >> $O/c12_box.txt
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) tools/encoder_jobs.py config4 > $O/c12_config4_n$n.json 2> $O/c12_config4_n$n.err
  echo "N=$n rc=$?"; cut -c1-420 $O/c12_config4_n$n.json
done
