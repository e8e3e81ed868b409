#!/bin/bash
# whole-encoder A/B on one GPU box: embed-stage trellis on the GPU (x264_pcamv) vs the reference's CPU stc_embed in the same
# host (x264_pcamv_cpustc, built with PCAMV_HOST_STC=1), 1080p, one stream; bitstreams must agree
cd $GRAFT_REPO_ROOT
./build/pcamv_synth 1920 1080 16 2 0 /tmp/c16.yuv 32
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
for b in x264_pcamv_cpustc x264_pcamv x264_pcamv_cpustc x264_pcamv; do
  echo "== $b"
  ( time PCAMV_STATS=/tmp/s.json ./host/_build/$b $A -o /tmp/$b.264 /tmp/c16.yuv 1920x1080 ) 2>&1 | grep -a -E "real|encoded"
  cat /tmp/s.json
done
md5sum /tmp/x264_pcamv_cpustc.264 /tmp/x264_pcamv.264
