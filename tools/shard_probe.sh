#!/bin/bash
# sharded whole-encoder parity (x264_pcamv --shards N --shard-frames K vs concatenated reference runs --seek/--frames)
cd $GRAFT_REPO_ROOT
./build/pcamv_synth 352 288 24 1 21 /tmp/sh.yuv 8
while IFS= read -r A; do
  [ -z "$A" ] && continue
  : > /tmp/want.264
  for g in 0 1 2 3 4 5; do ./oracle/_ref/x264_wide $A --seek $((g*4)) --frames 4 -o /tmp/rg.264 /tmp/sh.yuv 352x288 >/dev/null 2>&1; cat /tmp/rg.264 >> /tmp/want.264; done
  ./host/_build/x264_pcamv --shards 6 --shard-frames 4 $A -o /tmp/got.264 /tmp/sh.yuv 352x288 >/tmp/g.log 2>&1; r=$?
  [ "$(md5sum < /tmp/want.264)" = "$(md5sum < /tmp/got.264)" ] && s=OK || s="DIFF rc=$r $(grep -a -m1 pcamv /tmp/g.log | cut -c1-100)"
  echo "$s | shards 6x4 | $A"
done <<'LIST'
--qp 40 --ref 2 --keyint 250 --me umh --subme 5 --partitions all --emrate 0.3
--qp 46 --ref 1 --keyint 250 --me hex --subme 4 --emrate 0.2
--qp 30 --ref 2 --keyint 250 --me tesa --merange 8 --subme 5 --emrate 0.2
--qp 28 --ref 3 --keyint 250 --me esa --merange 8 --subme 3 --no-cabac --emrate 0.5
LIST
