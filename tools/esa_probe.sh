cd $GRAFT_REPO_ROOT
./build/pcamv_synth 1920 1080 3 3 0 /tmp/c3.yuv 32
A="--qp 26 --ref 4 --keyint 250 --me esa --merange 32 --subme 5 --emrate 0.2"
( time timeout 600 ./oracle/_ref/x264_wide $A -o /tmp/r3.264 /tmp/c3.yuv 1920x1080 ) 2>&1 | grep -a -E "real|encoded"
( time PCAMV_STATS=/tmp/s3.json timeout 900 ./host/_build/x264_pcamv $A -o /tmp/g3.264 /tmp/c3.yuv 1920x1080 ) 2>&1 | grep -a -E "real|encoded|pcamv"
cat /tmp/s3.json; md5sum /tmp/r3.264 /tmp/g3.264
