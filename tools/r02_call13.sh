#!/bin/bash
# round 2, GPU call 13: kind-affine scheduling of the split wavefront's control teams (phase length, patience, rows per team)
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt
timeout 300 python -m pytest tests/test_gpu_frame.py -m gpu -x -q -k "split" > $O/c13_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/c13_tests.log
PCAMV_SPLIT_STATS=1 PCAMV_QT_SWEEP="48:8:0,48:8:20000:3000,48:16:20000:3000,48:16:50000:5000,40:16:10000:2000,32:16:20000:5000" timeout 1200 python tools/quick_time.py 128 -2 1 > $O/c13_qt_split.log 2>&1; echo "qt rc=$?"
grep -v "^$" $O/c13_qt_split.log | cut -c1-420
