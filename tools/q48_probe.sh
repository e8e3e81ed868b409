cd $GRAFT_REPO_ROOT
./build/pcamv_synth 352 288 6 1 3 /tmp/o.yuv 24
A="--qp 48 --ref 2 --keyint 250 --me umh --subme 4 --emrate 0.2"
./oracle/_ref/x264_wide $A -o /tmp/r.264 /tmp/o.yuv 352x288 >/dev/null 2>&1
./host/_build/x264_pcamv $A -o /tmp/g.264 /tmp/o.yuv 352x288 >/dev/null 2>&1
PCAMV_PASS2_FULL=1 ./host/_build/x264_pcamv $A -o /tmp/f.264 /tmp/o.yuv 352x288 >/dev/null 2>&1
md5sum /tmp/r.264 /tmp/g.264 /tmp/f.264; cmp /tmp/r.264 /tmp/g.264 | head -2; ls -la /tmp/r.264 /tmp/g.264
for F in 2 3 4; do
  ./oracle/_ref/x264_wide $A --frames $F -o /tmp/r$F.264 /tmp/o.yuv 352x288 >/dev/null 2>&1
  ./host/_build/x264_pcamv $A --frames $F -o /tmp/g$F.264 /tmp/o.yuv 352x288 >/dev/null 2>&1
  echo "frames $F: $(md5sum < /tmp/r$F.264 | cut -c1-8) $(md5sum < /tmp/g$F.264 | cut -c1-8)"
done
