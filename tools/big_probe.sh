#!/bin/bash
cd $GRAFT_REPO_ROOT
./build/pcamv_synth 1920 1080 3 2 5 /tmp/b.yuv 32
while IFS= read -r A; do
  [ -z "$A" ] && continue
  ./oracle/_ref/x264_wide $A -o /tmp/r.264 /tmp/b.yuv 1920x1080 >/dev/null 2>&1; r1=$?
  ./host/_build/x264_pcamv $A -o /tmp/g.264 /tmp/b.yuv 1920x1080 >/tmp/g.log 2>&1; r2=$?
  [ "$(md5sum < /tmp/r.264)" = "$(md5sum < /tmp/g.264)" ] && [ $r1 = $r2 ] && s=OK || s="DIFF rc=$r1/$r2 $(grep -a -m1 pcamv /tmp/g.log | cut -c1-100)"
  echo "$s | 1080p | $A | $(grep -a -o 'encoded [0-9]* frames, [0-9.]* fps' /tmp/g.log)"
done <<'LIST'
--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --partitions all --emrate 0.2
--qp 30 --ref 2 --keyint 250 --me tesa --merange 16 --subme 5 --emrate 0.2
--qp 44 --ref 1 --keyint 250 --me hex --subme 5 --emrate 0.2
LIST
