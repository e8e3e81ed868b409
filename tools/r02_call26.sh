#!/bin/bash
# round 2, GPU call 26 (1 GPU): half-pel planes of device-built reference frames come back from the GPU (host skips x264_frame_filter)
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
timeout 900 python -m pytest tests/test_gpu_recon.py -m gpu -q -x -k "host_with" > $O/c26_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/c26_tests.log | cut -c1-400
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
print(refrun.synth_clip(pcamv, 1920, 1080, 40, config=2, stream=1, workdir='/dev/shm'))
PY
C=/dev/shm/clip_1920x1080_40_2_1_32.yuv
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
for rep in 1 2; do
for mode in default recon; do
  E=""; [ $mode = recon ] && E="PCAMV_DEVICE_RECON=1"
  env $E PCAMV_STATS=$O/c26_stats_$mode.json host/_build/x264_pcamv $A -o /dev/shm/o_$mode.264 $C 1920x1080 2>&1 | tail -1
  echo "$mode: $(cat $O/c26_stats_$mode.json)"
done; done
cmp /dev/shm/o_default.264 /dev/shm/o_recon.264 && echo same
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs PCAMV_JOB_DIGESTS=$PWD/profiles/r02_reference_digests
for mode in default recon default recon; do
  E=""; [ $mode = recon ] && E="PCAMV_DEVICE_RECON=1"
  env $E timeout 600 python tools/encoder_jobs.py config2 2>> $O/c26.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('config2 $mode: wall %.1f fps, loop %.1f fps, identical %s %s, per frame %s' % (d['encode_embed_fps'], d['encode_loop_fps'], d['bitstream_identical'], d['payload_identical'], {k: round(v, 1) for k, v in d['encoder_thread_ms_per_p_frame_rank0'].items()}))" | tee -a $O/c26_jobs.txt
done
