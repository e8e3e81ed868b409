#!/bin/bash
# end-of-round validation, second pass (device-built reference frames are the host's default now): GPU tests, smoke, both bench arms, one stream
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/final2_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/final2_tests.log; tail -4 $O/final2_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $O/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final2_smoke.log | cut -c1-300
python bench.py > $O/r02_bench_final_s128.json 2> $O/final2_bench.err; echo "bench rc=$?"; tail -c 300 $O/final2_bench.err; wc -l $O/r02_bench_final_s128.json; cut -c1-200 $O/r02_bench_final_s128.json
python bench.py --impl reference > $O/r02_bench_final_reference_arm.json 2> $O/final2_ref.err; echo "ref arm rc=$?"; wc -l $O/r02_bench_final_reference_arm.json
C=/dev/shm/clip40.yuv
python - <<'PY'
import sys, shutil; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import pcamv_loader, refrun
pcamv = pcamv_loader.load()
shutil.move(refrun.synth_clip(pcamv, 1920, 1080, 40, config=2, stream=1, workdir='/dev/shm'), '/dev/shm/clip40.yuv')
PY
A="--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
PCAMV_STATS=$O/final2_stream_stats.json host/_build/x264_pcamv $A -o /dev/shm/o.264 $C 1920x1080 2>&1 | tail -1 | tee $O/final2_stream.txt; cat $O/final2_stream_stats.json
