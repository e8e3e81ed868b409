#!/bin/bash
# BASELINE config 5 (and config 2) as embed + extract round trips through the .264 alone, conformant mode; the conformant GPU tests with the rebuilt host
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
mkdir -p $O
T0=$(date +%s)
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
timeout 85 python tools/round_trip_job.py config5 > $O/r02_round_trip_config5.json 2> $O/c31_rt5.err; echo "config5 rc=$? t=$(( $(date +%s) - T0 ))"; cut -c1-900 $O/r02_round_trip_config5.json; tail -3 $O/c31_rt5.err | cut -c1-300
timeout 40 python -m pytest tests/test_z_gpu_conformant.py -q -n 4 > $O/c31_conformant.log 2>&1; echo "conformant rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 $O/c31_conformant.log | cut -c1-200
rm -rf /dev/shm/pcamv_jobs/config5
timeout 45 python tools/round_trip_job.py config2 > $O/r02_round_trip_config2.json 2> $O/c31_rt2.err; echo "config2 rc=$? t=$(( $(date +%s) - T0 ))"; cut -c1-900 $O/r02_round_trip_config2.json; tail -3 $O/c31_rt2.err | cut -c1-300
