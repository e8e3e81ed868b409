#!/bin/bash
# whole-encoder md5 parity over sizes / seeds / noise levels with two option sets
cd $GRAFT_REPO_ROOT
for spec in "64 64 8 7 0" "128 96 8 8 4" "80 48 10 9 16" "720 480 5 10 32" "176 144 12 11 2" "352 288 10 12 0" "640 368 5 13 64" "32 16 8 14 8" "96 80 9 15 1"; do
  set -- $spec
  ./build/pcamv_synth $1 $2 $3 1 $4 /tmp/s.yuv $5
  for A in "--qp 30 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.25" "--qp 40 --ref 1 --keyint 4 --min-keyint 4 --me hex --subme 4 --partitions all --emrate 0.4"; do
    ./oracle/_ref/x264_wide $A -o /tmp/r.264 /tmp/s.yuv $1x$2 >/dev/null 2>&1; r1=$?
    ./host/_build/x264_pcamv $A -o /tmp/g.264 /tmp/s.yuv $1x$2 >/tmp/g.log 2>&1; r2=$?
    a=$(md5sum < /tmp/r.264 | cut -c1-8); b=$(md5sum < /tmp/g.264 | cut -c1-8)
    [ "$a" = "$b" ] && [ $r1 = $r2 ] && s=OK || s="DIFF rc=$r1/$r2 $(grep -a -m1 pcamv /tmp/g.log | cut -c1-100)"
    echo "$s | $1x$2 f$3 seed$4 noise$5 | $A"
  done
done
