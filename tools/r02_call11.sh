#!/bin/bash
# round 2, GPU call 11 (1 GPU): GPU parity tests under the PCAMV_CHECKED library + its negative control; one-context numbers;
# BASELINE config 4 at full size on one GPU
cd $GRAFT_REPO_ROOT
O=gpurun_out
L=$PWD/build/variants/lib_checked.so
( echo "# PCAMV_LIB=build/variants/lib_checked.so (tools/checked_build.sh): GPU parity tests, then the negative control";
  PCAMV_LIB=$L timeout 1200 python -m pytest tests/test_gpu_frame.py tests/test_gpu_kernels.py tests/test_gpu_embed.py -m gpu -q 2>&1 | tail -4;
  PCAMV_LIB=$L timeout 300 python tools/probes/checked_probe.py 2>&1 | tail -3 ) > $O/c11_checked.txt 2>&1; cat $O/c11_checked.txt
timeout 600 python bench.py --streams 1 --steps 5 --warmup 3 --no-cpu-baseline --encoder-job none > $O/c11_bench_s1.json 2> $O/c11_bench_s1.err; echo "s1 rc=$?"
timeout 600 python bench.py --streams 1 --steps 5 --warmup 3 --no-cpu-baseline --encoder-job none --pass2-elide > $O/c11_bench_s1_elide.json 2>> $O/c11_bench_s1.err; echo "s1 elide rc=$?"
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
timeout 1500 python tools/encoder_jobs.py config4 > $O/c11_config4_n1.json 2> $O/c11_config4_n1.err; echo "config4 rc=$?"; cut -c1-700 $O/c11_config4_n1.json; tail -c 300 $O/c11_config4_n1.err
