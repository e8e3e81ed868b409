#!/bin/bash
# round 2, GPU call 18 (1 GPU): rows behind the wavefront (pcamv_analyse_p_begin / _rows) — API parity test, whole-encoder parity
# tests, single-stream A/B, BASELINE config 2 / 5 jobs
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
timeout 1500 python -m pytest tests/test_gpu_frame.py tests/test_gpu_host.py tests/test_gpu_recon.py tests/test_extract.py -m gpu -q -x > $O/c18_tests.log 2>&1; echo "tests rc=$?"; tail -12 $O/c18_tests.log | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
print(refrun.synth_clip(pcamv, 1920, 1080, 40, config=2, stream=1, workdir='/dev/shm'))
PY
C=/dev/shm/clip_1920x1080_40_2_1_32.yuv
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
for rep in 1 2; do
for mode in stream nostream; do
  E=""; [ $mode = nostream ] && E="PCAMV_NO_ROW_STREAM=1"
  env $E PCAMV_STATS=$O/c18_stats_$mode.json host/_build/x264_pcamv $A -o /dev/shm/o_$mode.264 $C 1920x1080 2>&1 | tail -1
  echo "$mode: $(cat $O/c18_stats_$mode.json)"
done; done
cmp /dev/shm/o_stream.264 /dev/shm/o_nostream.264 && echo same
env PCAMV_ROWS_PER_CTA=4 PCAMV_STATS=$O/c18_stats_rpc4.json host/_build/x264_pcamv $A -o /dev/shm/o_rpc4.264 $C 1920x1080 2>&1 | tail -1; echo "rpc4: $(cat $O/c18_stats_rpc4.json)"
env PCAMV_DEVICE_RECON=1 PCAMV_STATS=$O/c18_stats_recon.json host/_build/x264_pcamv $A -o /dev/shm/o_recon.264 $C 1920x1080 2>&1 | tail -1; echo "recon: $(cat $O/c18_stats_recon.json)"; cmp /dev/shm/o_stream.264 /dev/shm/o_recon.264 && echo same
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
timeout 900 python tools/encoder_jobs.py config2 config5 > $O/c18_jobs.json 2> $O/c18_jobs.err; echo "jobs rc=$?"; cut -c1-700 $O/c18_jobs.json; tail -c 300 $O/c18_jobs.err
PCAMV_NO_ROW_STREAM=1 timeout 900 python tools/encoder_jobs.py config2 > $O/c18_jobs_nostream.json 2>> $O/c18_jobs.err; cut -c1-700 $O/c18_jobs_nostream.json
