cd $GRAFT_REPO_ROOT
for spec in "48 32 6" "16 16 5" "352 288 1" "1024 16 4" "16 512 4"; do
  set -- $spec
  ./build/pcamv_synth $1 $2 $3 1 0 /tmp/e.yuv 32
  A="--qp 26 --ref 2 --keyint 250 --me umh --subme 5 --emrate 0.3"
  ./oracle/_ref/x264_wide $A -o /tmp/r.264 /tmp/e.yuv $1x$2 >/dev/null 2>&1; r1=$?
  ./host/_build/x264_pcamv $A -o /tmp/g.264 /tmp/e.yuv $1x$2 >/tmp/g.log 2>&1; r2=$?
  echo "$spec ref_rc=$r1 gpu_rc=$r2 $(md5sum < /tmp/r.264 | cut -c1-8) $(md5sum < /tmp/g.264 | cut -c1-8) $(tail -c 200 /tmp/g.log | tr '\n' ' ' | cut -c1-150)"
done
