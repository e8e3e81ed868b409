#!/bin/bash
# round 2, GPU call 5: candidate-parallel cost-table kernel (parity tests, then A/B against the one-team kernel and the pre-refactor library)
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt
df -h /tmp /dev/shm > $O/c5_df.txt 2>&1; nproc >> $O/c5_df.txt; free -g >> $O/c5_df.txt
timeout 900 python -m pytest tests/test_gpu_frame.py -m gpu -x -q > $O/c5_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c5_tests.log
( tools/ab.sh "lib_pre.so 128 4" "default 128 4"; PCAMV_COST_TABLE=team tools/ab.sh "default 128 4"; tools/ab.sh "default 128 4" ) > $O/c5_ab.log 2>&1
cat $O/c5_ab.log
