#!/bin/bash
# round 2, GPU call 19 (1 GPU): rendezvous groups x wavefront layout of the whole-encoder jobs
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
: > $O/c19_sweep.txt
for job in config2 config5; do
  for rpc in 4 1; do
    for g in 2 4 8 16 64; do
      [ $job = config2 ] && [ $g = 64 ] && continue
      PCAMV_JOB_GROUPS=$g PCAMV_JOB_RPC=$rpc timeout 600 python tools/encoder_jobs.py $job 2>> $O/c19.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$job groups $g rpc $rpc: wall %.1f fps, loop %.1f fps, gpu share %.2f, identical %s %s' % (d['encode_embed_fps'], d['encode_loop_fps'], d['gpu_call_share_of_encoder_thread'], d['bitstream_identical'], d['payload_identical']))" | tee -a $O/c19_sweep.txt
    done
  done
done
tail -3 $O/c19.err
