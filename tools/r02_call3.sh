#!/bin/bash
# round 2, GPU call 3: where does the split wavefront spend its time (team timers), and did the resumable refactor cost the row-group kernel anything
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt
PCAMV_SPLIT_STATS=1 PCAMV_QT_SWEEP="48:6,32:6" timeout 900 python tools/quick_time.py 128 -2 1 > $O/c3_qt_split.log 2>&1; echo "qt split rc=$?"; grep -v "^$" $O/c3_qt_split.log | tail -12
tools/ab.sh "default 128 4" "lib_pre.so 128 4" "default 128 4" "lib_pre.so 128 4" > $O/c3_ab.log 2>&1; cat $O/c3_ab.log
