#!/bin/bash
# round 2, GPU call 22 (2 GPUs): bench.py at N = 2 (weak-scaling frame seam + the sharded encoder job with the NCCL gather in the
# clock), BASELINE config 4 over 2 GPUs (parity against the committed reference digests)
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
nproc > $O/c22_box.txt; nvidia-smi -L >> $O/c22_box.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 6 --warmup 3 > $O/r02_bench_final_2gpu.json 2> $O/c22_bench.err; echo "bench N=2 rc=$?"; cut -c1-300 $O/r02_bench_final_2gpu.json; tail -c 300 $O/c22_bench.err
export PCAMV_JOB_DIGESTS=$PWD/profiles/r02_reference_digests
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 tools/encoder_jobs.py config4 config2 > $O/c22_jobs_n2.json 2> $O/c22_jobs.err; echo "jobs N=2 rc=$?"; cut -c1-500 $O/c22_jobs_n2.json; tail -c 300 $O/c22_jobs.err
