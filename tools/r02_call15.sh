#!/bin/bash
# round 2, GPU call 15 (1 GPU): pass 1 without the host (default) — the whole GPU suite, then BASELINE configs 2 and 5 through
# the whole encoder with bitstream + payload parity
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/c15_tests.log 2>&1; echo "tests rc=$?"; tail -25 $O/c15_tests.log | cut -c1-400
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
timeout 900 python tools/encoder_jobs.py config2 config5 > $O/c15_jobs.json 2> $O/c15_jobs.err; echo "jobs rc=$?"; cut -c1-900 $O/c15_jobs.json; tail -c 400 $O/c15_jobs.err
