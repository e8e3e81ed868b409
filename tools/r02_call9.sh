#!/bin/bash
# round 2, GPU call 9 (1 GPU): BASELINE configs 2 and 5 at their stated sizes through the whole encoder; ncu launch list of the
# bench command and full captures of every kernel at HEAD
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt PCAMV_JOB_DIR=/dev/shm/pcamv_jobs
timeout 1200 python tools/encoder_jobs.py config2 config5 > $O/c9_jobs.json 2> $O/c9_jobs.err; echo "jobs rc=$?"; cat $O/c9_jobs.json | cut -c1-900; tail -c 600 $O/c9_jobs.err
rm -rf /dev/shm/pcamv_jobs
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encoder-job none > $O/c9_b_short.json 2> $O/c9_b_short.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file $O/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encoder-job none > $O/c9_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 python tools/quick_time.py 128 4 1 > $O/c9_qt.log 2>&1
for k in k_analyse_p_batch k_cost_table_cta_batch; do
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/prof_r02_$k python tools/quick_time.py 128 4 1 > $O/c9_ncu_$k.log 2>&1; echo "$k rc=$?"
done
# the frame-level kernels of one context: border, half-pel filter, box sums (esa context), trellis, embed stage
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expand_border|k_hpel_filter|k_stc_forward|k_stc_backward|k_embed_' -c 12 -f -o $O/prof_r02_small \
    python -m pytest tests/test_gpu_embed.py -m gpu -x -q -k "golden and umh5" > $O/c9_ncu_small.log 2>&1; echo "small kernels rc=$?"
ls -la $O | tail -12
