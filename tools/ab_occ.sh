#!/bin/bash
export PCAMV_QT_DIR=/tmp/pcamv_qt
run() { # lib warps_per_sm rows_per_cta
  if [ "$1" = default ]; then unset PCAMV_LIB; else export PCAMV_LIB=$PWD/build/variants/$1; fi
  export PCAMV_BATCH_WARPS_PER_SM=$2
  echo -n "warps/SM $2 "; timeout 600 python tools/quick_time.py 128 $3 3 2>&1 | tail -1
}
run lib_c4.so 16 4; run lib_c4.so 16 -1; run lib_c5.so 20 4; run lib_c5.so 20 -1; run default 24 -1; run lib_c8.so 32 4; run lib_c8.so 32 -1
