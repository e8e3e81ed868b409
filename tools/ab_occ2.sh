#!/bin/bash
export PCAMV_QT_DIR=/tmp/pcamv_qt
unset PCAMV_LIB
for spec in "4 4" "8 4" "8 -1" "12 4" "12 -1" "16 -1" "24 4"; do
  set -- $spec
  export PCAMV_BATCH_WARPS_PER_SM=$1
  echo -n "warps/SM $1 "; timeout 600 python tools/quick_time.py 128 $2 2 2>&1 | tail -1
done
