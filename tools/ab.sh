#!/bin/bash
# A/B of library builds under build/variants on one GPU box: tools/ab.sh "<lib> <S> <rows_per_cta>" ...
export PCAMV_QT_DIR=/tmp/pcamv_qt
for spec in "$@"; do
  set -- $spec
  lib=$1; S=${2:-128}; rpc=${3:-4}
  if [ "$lib" = default ]; then unset PCAMV_LIB; else export PCAMV_LIB=$PWD/build/variants/$lib; fi
  timeout 600 python tools/quick_time.py $S $rpc 3 2>&1 | tail -2
done
