#!/bin/bash
# round 2, GPU call 20 (1 GPU): where the GPU-call time of an encoder thread goes (per-stage timers), one stream and config 2
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
print(refrun.synth_clip(pcamv, 1920, 1080, 40, config=2, stream=1, workdir='/dev/shm'))
PY
C=/dev/shm/clip_1920x1080_40_2_1_32.yuv
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
for rep in 1 2; do
  PCAMV_STATS=$O/c20_stats.json host/_build/x264_pcamv $A -o /dev/shm/o.264 $C 1920x1080 2>&1 | tail -1; cat $O/c20_stats.json
done
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs PCAMV_JOB_DIGESTS=$PWD/profiles/r02_reference_digests
timeout 600 python tools/encoder_jobs.py config2 config2 > $O/c20_jobs.json 2> $O/c20_jobs.err; echo "jobs rc=$?"; cut -c1-1100 $O/c20_jobs.json; tail -c 300 $O/c20_jobs.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
