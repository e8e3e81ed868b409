#!/bin/bash
# round 2, GPU call 4: control-code tokens per SM (row-group kernel), HEAD vs the pre-refactor library
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt
( tools/ab.sh "lib_pre.so 128 4" "default 128 4"
  for k in 1 2 3 4 6 8 12; do echo "tokens $k"; PCAMV_CTL_TOKENS=$k tools/ab.sh "default 128 4"; done ) > $O/c4_ab.log 2>&1
cat $O/c4_ab.log
