#!/bin/bash
# determinism: the same encode repeated must give the same bytes (a wavefront race would show up as a changing md5)
cd $GRAFT_REPO_ROOT
./build/pcamv_synth 1280 720 4 2 6 /tmp/d.yuv 16
./build/pcamv_synth 352 288 24 1 22 /tmp/ds.yuv 4
A="--qp 34 --ref 2 --keyint 250 --me umh --subme 5 --partitions all --emrate 0.2"
for i in 1 2 3 4 5 6; do ./host/_build/x264_pcamv $A -o /tmp/d$i.264 /tmp/d.yuv 1280x720 >/dev/null 2>&1; md5sum < /tmp/d$i.264; done | sort | uniq -c
for i in 1 2 3 4 5 6; do PCAMV_ROWS_PER_CTA=4 ./host/_build/x264_pcamv --shards 6 --shard-frames 4 $A -o /tmp/s$i.264 /tmp/ds.yuv 352x288 >/dev/null 2>&1; md5sum < /tmp/s$i.264; done | sort | uniq -c
for i in 1 2 3; do PCAMV_ROWS_PER_CTA=-1 ./host/_build/x264_pcamv --shards 6 --shard-frames 4 $A -o /tmp/p$i.264 /tmp/ds.yuv 352x288 >/dev/null 2>&1; md5sum < /tmp/p$i.264; done | sort | uniq -c
