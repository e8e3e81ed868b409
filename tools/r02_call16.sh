#!/bin/bash
# round 2, GPU call 16 (1 GPU): where a single-stream encode spends its host time (gprof twin of x264_pcamv), with and without
# the host's own pass 1
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
print(refrun.synth_clip(pcamv, 1920, 1080, 16, config=2, stream=1, workdir='/dev/shm'))
PY
C=/dev/shm/clip_1920x1080_16_2_1_32.yuv
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
mkdir -p /dev/shm/p1 /dev/shm/p2
( cd /dev/shm/p1; PCAMV_STATS=$O/c16_stats_direct.json $GRAFT_REPO_ROOT/host/_build/x264_pcamv_pg $A -o /dev/shm/p1/o.264 $C 1920x1080 2>&1 | tail -3; gprof -b -p $GRAFT_REPO_ROOT/host/_build/x264_pcamv_pg gmon.out 2>/dev/null | head -60 > $O/c16_gprof_direct.txt )
( cd /dev/shm/p2; PCAMV_HOST_PASS1=1 PCAMV_STATS=$O/c16_stats_hostpass1.json $GRAFT_REPO_ROOT/host/_build/x264_pcamv_pg $A -o /dev/shm/p2/o.264 $C 1920x1080 2>&1 | tail -3; gprof -b -p $GRAFT_REPO_ROOT/host/_build/x264_pcamv_pg gmon.out 2>/dev/null | head -60 > $O/c16_gprof_hostpass1.txt )
cmp /dev/shm/p1/o.264 /dev/shm/p2/o.264 && echo same
for m in direct hostpass1; do echo "== $m"; cat $O/c16_stats_$m.json; done
PCAMV_DEVICE_RECON=1 PCAMV_STATS=$O/c16_stats_recon.json host/_build/x264_pcamv $A -o /dev/shm/p1/o2.264 $C 1920x1080 2>&1 | tail -2; cmp /dev/shm/p1/o.264 /dev/shm/p1/o2.264 && echo same; cat $O/c16_stats_recon.json
PCAMV_STATS=$O/c16_stats_plain.json host/_build/x264_pcamv $A -o /dev/shm/p1/o3.264 $C 1920x1080 2>&1 | tail -2; cat $O/c16_stats_plain.json
head -40 $O/c16_gprof_direct.txt
