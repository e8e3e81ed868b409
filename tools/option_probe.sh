#!/bin/bash
# whole-encoder md5 parity over option sets that no test covers yet (CIF, 6 frames); prints one line per set
cd $GRAFT_REPO_ROOT
./build/pcamv_synth 352 288 ${PROBE_FRAMES:-8} 1 ${PROBE_STREAM:-3} /tmp/o.yuv ${PROBE_NOISE:-8}
while IFS= read -r A; do
  [ -z "$A" ] && continue
  ./oracle/_ref/x264_wide $A -o /tmp/r.264 /tmp/o.yuv 352x288 >/dev/null 2>&1; r1=$?
  ./host/_build/x264_pcamv $A -o /tmp/g.264 /tmp/o.yuv 352x288 >/tmp/g.log 2>&1; r2=$?
  a=$(md5sum < /tmp/r.264 | cut -c1-8); b=$(md5sum < /tmp/g.264 | cut -c1-8)
  [ "$a" = "$b" ] && [ $r1 = $r2 ] && s=OK || s="DIFF rc=$r1/$r2 $(grep -a -m1 pcamv /tmp/g.log | cut -c1-120)"
  echo "$s | $A"
done <<'LIST'
--qp 48 --ref 2 --keyint 250 --me umh --subme 4 --emrate 0.2
--qp 40 --ref 1 --keyint 250 --me hex --subme 5 --emrate 0.2
--qp 44 --ref 3 --keyint 250 --me umh --subme 5 --partitions all --emrate 0.3
--qp 51 --ref 1 --keyint 250 --me dia --subme 3 --emrate 0.5
--qp 38 --ref 2 --keyint 250 --me esa --merange 8 --subme 5 --no-cabac --emrate 0.2
--qp 42 --ref 1 --keyint 250 --me tesa --merange 8 --subme 5 --partitions p8x8,p4x4 --emrate 0.2
--qp 36 --ref 4 --keyint 4 --min-keyint 4 --me hex --subme 4 --no-fast-pskip --emrate 0.2
--qp 45 --ref 1 --keyint 250 --me hex --subme 5 --no-dct-decimate --emrate 0.2
--qp 30 --ref 2 --keyint 250 --me umh --subme 1 --emrate 0.2
--qp 46 --ref 2 --keyint 250 --me hex --subme 3 --partitions none --emrate 0.9
LIST
