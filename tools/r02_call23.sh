#!/bin/bash
# round 2, GPU call 23 (4 GPUs): BASELINE config 4 and config 5 over 4 GPUs (parity against the committed reference digests)
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs PCAMV_JOB_DIGESTS=$PWD/profiles/r02_reference_digests
nproc > $O/c23_box.txt; nvidia-smi -L >> $O/c23_box.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29704 tools/encoder_jobs.py config4 config5 > $O/c23_jobs_n4.json 2> $O/c23_jobs.err; echo "jobs N=4 rc=$?"; cut -c1-500 $O/c23_jobs_n4.json; tail -c 300 $O/c23_jobs.err
