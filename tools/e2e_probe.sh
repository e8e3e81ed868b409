cd $GRAFT_REPO_ROOT
./build/pcamv_synth 1920 1080 24 2 0 /tmp/c24.yuv 32
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
for v in "" "PCAMV_NO_PINNED=1"; do
echo "== $v"
( time env $v PCAMV_STATS=/tmp/s.json ./host/_build/x264_pcamv $A -o /tmp/g.264 /tmp/c24.yuv 1920x1080 ) 2>&1 | grep -a -E "real|encoded"
cat /tmp/s.json
done
