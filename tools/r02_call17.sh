#!/bin/bash
# round 2, GPU call 17 (1 GPU): stats-only intra analysis skipped on the host — whole-encoder parity tests, gprof of a single
# stream, start-up probe, BASELINE config 2
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
timeout 1500 python -m pytest tests/test_gpu_host.py tests/test_gpu_recon.py tests/test_extract.py -m gpu -q > $O/c17_tests.log 2>&1; echo "tests rc=$?"; tail -8 $O/c17_tests.log | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
print(refrun.synth_clip(pcamv, 1920, 1080, 16, config=2, stream=1, workdir='/dev/shm'))
PY
C=/dev/shm/clip_1920x1080_16_2_1_32.yuv
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
mkdir -p /dev/shm/p1
( cd /dev/shm/p1; PCAMV_STATS=$O/c17_stats_pg.json $GRAFT_REPO_ROOT/host/_build/x264_pcamv_pg $A -o /dev/shm/p1/o.264 $C 1920x1080 2>&1 | tail -1; gprof -b -p $GRAFT_REPO_ROOT/host/_build/x264_pcamv_pg gmon.out 2>/dev/null | head -40 > $O/c17_gprof.txt )
PCAMV_STATS=$O/c17_stats.json host/_build/x264_pcamv $A -o /dev/shm/p1/o3.264 $C 1920x1080 2>&1 | tail -1; cat $O/c17_stats.json
PCAMV_HOST_INTRA=1 PCAMV_STATS=$O/c17_stats_intra.json host/_build/x264_pcamv $A -o /dev/shm/p1/o4.264 $C 1920x1080 2>&1 | tail -1; cat $O/c17_stats_intra.json; cmp /dev/shm/p1/o3.264 /dev/shm/p1/o4.264 && echo same
head -30 $O/c17_gprof.txt
python tools/probes/open_probe.py 2>&1 | tee $O/c17_open_probe.txt
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
timeout 900 python tools/encoder_jobs.py config2 > $O/c17_jobs.json 2> $O/c17_jobs.err; echo "jobs rc=$?"; cut -c1-420 $O/c17_jobs.json; tail -c 300 $O/c17_jobs.err
