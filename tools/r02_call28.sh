#!/bin/bash
# round 2, GPU call 28 (1 GPU): the two cases of the random sweep where check mode reported a GPU-built reference picture that
# differs from the host's
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
S=build/pcamv_synth
$S 48 32 6 1 271 /dev/shm/b.yuv 0; $S 16 96 6 1 219 /dev/shm/a.yuv 0
( PCAMV_CHECK_RECON=1 PCAMV_STATS=$O/c28_b.json host/_build/x264_pcamv --qp 26 --ref 4 --keyint 250 --me hex --merange 8 --subme 1 --keyint 3 --min-keyint 3 -o /dev/shm/b.264 /dev/shm/b.yuv 48x32 2>&1 | grep -a "pcamv" ; cat $O/c28_b.json
  PCAMV_CHECK_RECON=1 PCAMV_STATS=$O/c28_a.json host/_build/x264_pcamv --qp 38 --ref 4 --keyint 250 --me tesa --merange 12 --subme 2 --emrate 0.7 -o /dev/shm/a.264 /dev/shm/a.yuv 16x96 2>&1 | grep -a "pcamv"; cat $O/c28_a.json
  oracle/_ref/x264_wide --qp 26 --ref 4 --keyint 250 --me hex --merange 8 --subme 1 --keyint 3 --min-keyint 3 -o /dev/shm/b_ref.264 /dev/shm/b.yuv 48x32 > /dev/null 2>&1; cmp /dev/shm/b.264 /dev/shm/b_ref.264 && echo "b: same bitstream"
  oracle/_ref/x264_wide --qp 38 --ref 4 --keyint 250 --me tesa --merange 12 --subme 2 --emrate 0.7 -o /dev/shm/a_ref.264 /dev/shm/a.yuv 16x96 > /dev/null 2>&1; cmp /dev/shm/a.264 /dev/shm/a_ref.264 && echo "a: same bitstream"
  PCAMV_STATS=$O/c28_b2.json host/_build/x264_pcamv --qp 26 --ref 4 --keyint 250 --me hex --merange 8 --subme 1 --keyint 3 --min-keyint 3 -o /dev/shm/b2.264 /dev/shm/b.yuv 48x32 > /dev/null 2>&1; cmp /dev/shm/b2.264 /dev/shm/b_ref.264 && echo "b default mode: same bitstream"
  PCAMV_STATS=$O/c28_a2.json host/_build/x264_pcamv --qp 38 --ref 4 --keyint 250 --me tesa --merange 12 --subme 2 --emrate 0.7 -o /dev/shm/a2.264 /dev/shm/a.yuv 16x96 > /dev/null 2>&1; cmp /dev/shm/a2.264 /dev/shm/a_ref.264 && echo "a default mode: same bitstream"
) 2>&1 | tee $O/c28.txt | cut -c1-400
