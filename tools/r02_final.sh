#!/bin/bash
# end-of-round validation on one GPU box: GPU tests, smoke, both bench arms, one stream, ncu of the reconstruction kernels,
# BASELINE config 4 at full size (reference digests: profiles/r02_reference_digests)
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/final_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/final_tests.log; tail -4 $O/final_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log | cut -c1-300
python bench.py --impl reference > $O/r02_bench_final_reference_arm.json 2> $O/final_ref.err; echo "ref arm rc=$?"
python bench.py > $O/r02_bench_final_s128.json 2> $O/final_bench.err; echo "bench rc=$?"; tail -c 400 $O/final_bench.err; cut -c1-600 $O/r02_bench_final_s128.json
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
print(refrun.synth_clip(pcamv, 1920, 1080, 40, config=2, stream=1, workdir='/dev/shm'))
PY
C=/dev/shm/clip_1920x1080_40_2_1_32.yuv
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
PCAMV_STATS=$O/final_stream_stats.json host/_build/x264_pcamv $A -o /dev/shm/o.264 $C 1920x1080 2>&1 | tail -1 | tee $O/final_stream.txt; cat $O/final_stream_stats.json
PCAMV_DEVICE_RECON=1 host/_build/x264_pcamv $A --frames 4 -o /dev/shm/o2.264 $C 1920x1080 > /dev/null 2>&1 && \
PCAMV_DEVICE_RECON=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_recon|k_deblock' -c 6 -f -o $O/prof_r02_recon \
    host/_build/x264_pcamv $A --frames 4 -o /dev/shm/o2.264 $C 1920x1080 > $O/final_ncu_recon.log 2>&1; echo "ncu recon rc=$?"
export PCAMV_JOB_DIR=/dev/shm/pcamv_jobs PCAMV_JOB_DIGESTS=$PWD/profiles/r02_reference_digests
rm -rf /dev/shm/pcamv_jobs
timeout 900 python tools/encoder_jobs.py config4 config2 config5 > $O/final_jobs.json 2> $O/final_jobs.err; echo "jobs rc=$?"; cut -c1-500 $O/final_jobs.json; tail -c 300 $O/final_jobs.err
