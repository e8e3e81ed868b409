#!/bin/bash
# round 2, GPU call 25 (1 GPU): boundary strengths of the frame computed ahead of the deblocking wavefront — parity tests, kernel time
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
timeout 900 python -m pytest tests/test_gpu_recon.py tests/test_gpu_host.py -m gpu -q -x -k "recon or switches" > $O/c25_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/c25_tests.log | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
print(refrun.synth_clip(pcamv, 1920, 1080, 40, config=2, stream=1, workdir='/dev/shm'))
PY
C=/dev/shm/clip_1920x1080_40_2_1_32.yuv
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
PCAMV_DEVICE_RECON=1 PCAMV_CHECK_RECON=1 PCAMV_STATS=$O/c25_stats_recon.json host/_build/x264_pcamv $A -o /dev/shm/o.264 $C 1920x1080 2>&1 | tail -1; cat $O/c25_stats_recon.json
PCAMV_DEVICE_RECON=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_recon|k_deblock' -c 6 -f -o $O/prof_r02_recon_v3 \
    host/_build/x264_pcamv $A --frames 4 -o /dev/shm/o2.264 $C 1920x1080 > $O/c25_ncu_recon.log 2>&1; echo "ncu recon rc=$?"
