cd $GRAFT_REPO_ROOT
N=${1:-16}; K=${2:-12}; export PCAMV_GROUPS=${3:-2}
[ -f /tmp/cs_$((N*K)).yuv ] || ./build/pcamv_synth 1920 1080 $((N*K)) 2 0 /tmp/cs_$((N*K)).yuv 32
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
S=$(date +%s.%N)
PCAMV_VERBOSE=1 PCAMV_ROWS_PER_CTA=4 timeout 120 ./host/_build/x264_pcamv --shards $N --shard-frames $K $A -o /tmp/gs.264 /tmp/cs_$((N*K)).yuv 1920x1080 > /tmp/sp.out 2> /tmp/sp.err
RC=$?
E=$(date +%s.%N)
grep -a "encoded" /tmp/sp.err | awk -v rc=$RC -v n=$N -v k=$K -v g=$PCAMV_GROUPS -v w=$(echo "$E - $S" | bc) '{s+=$5; c++} END{printf "shards %d x %d frames, groups %d: rc %d, %d finished, sum of loop fps %.1f, wall %.2f s -> %.1f fps\n", n, k, g, rc, c, s, w, n*k/w}'
grep -a "pcamv\]" /tmp/sp.err | tail -2
