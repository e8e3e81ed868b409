set -e
cd $GRAFT_REPO_ROOT
./build/pcamv_synth 1920 1080 24 2 0 /tmp/c24.yuv 32
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
( time PCAMV_STATS=/tmp/s.json ./host/_build/x264_pcamv $A -o /tmp/g.264 /tmp/c24.yuv 1920x1080 ) 2>&1 | grep -a -E "real|encoded" 
cat /tmp/s.json
( time ./oracle/_ref/x264_wide $A -o /tmp/r.264 /tmp/c24.yuv 1920x1080 ) 2>&1 | grep -a -E "real|encoded"
md5sum /tmp/g.264 /tmp/r.264
( time PCAMV_STATS=/tmp/s.json ./host/_build/x264_pcamv $A --frames 1 -o /tmp/g1.264 /tmp/c24.yuv 1920x1080 ) 2>&1 | grep -a -E "real|encoded"
