#!/bin/bash
# round 2, GPU call 2: split wavefront — parity on the small batch tests, then 1080p x 128 contexts timing over a sweep of
# (control SMs : row slots per control team)
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt
timeout 600 python -m pytest tests/test_gpu_frame.py -m gpu -x -q -k "batch" > $O/c2_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/c2_tests.log
PCAMV_QT_SWEEP="32:6,24:6,40:6,48:6,32:4,32:8,56:8" timeout 900 python tools/quick_time.py 128 -2 2 > $O/c2_qt_split.log 2>&1; echo "qt split rc=$?"; tail -8 $O/c2_qt_split.log
timeout 600 python tools/quick_time.py 128 4 2 > $O/c2_qt_base.log 2>&1; echo "qt base rc=$?"; tail -1 $O/c2_qt_base.log
