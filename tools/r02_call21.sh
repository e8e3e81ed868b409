#!/bin/bash
# round 2, GPU call 21 (1 GPU): host threads waiting for the GPU sleep (PCAMV_BLOCKING_SYNC=1) or spin (0) — config 2 / config 5 jobs
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs PCAMV_JOB_DIGESTS=$PWD/profiles/r02_reference_digests
: > $O/c21_sweep.txt
for rep in 1 2; do
for job in config2 config5; do
  for bs in 0 1; do
      PCAMV_BLOCKING_SYNC=$bs timeout 600 python tools/encoder_jobs.py $job 2>> $O/c21.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$job blocking_sync $bs: wall %.1f fps, loop %.1f fps, gpu share %.2f, identical %s %s, per frame %s' % (d['encode_embed_fps'], d['encode_loop_fps'], d['gpu_call_share_of_encoder_thread'], d['bitstream_identical'], d['payload_identical'], {k: round(v, 1) for k, v in d['encoder_thread_ms_per_p_frame_rank0'].items()}))" | tee -a $O/c21_sweep.txt
  done
done; done
tail -3 $O/c21.err
