#!/bin/bash
# end-of-round validation on one GPU box: GPU tests, smoke, both bench arms, then (each only after its plain run exited 0)
# the ncu launch list of the bench command and one full capture of the dominant kernel at the bench's 128 contexts
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/final_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/final_tests.log; tail -3 $O/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log
python bench.py --impl reference > $O/bench_r01_v2_ref.json 2> $O/bench_r01_v2_ref.err; echo "ref arm rc=$?"
python bench.py > $O/bench_r01_v2.json 2> $O/bench_r01_v2.err; echo "bench rc=$?"; tail -c 600 $O/bench_r01_v2.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/b_short.json 2> $O/b_short.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $O/launches_r01_v2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
PCAMV_QT_DIR=/tmp/qt python tools/quick_time.py 128 4 2 > $O/qt.log 2>&1 && \
PCAMV_QT_DIR=/tmp/qt timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_analyse_p_batch -s 1 -c 1 -f -o $O/prof_batch128_r01_v2 \
    python tools/quick_time.py 128 4 1 > $O/ncu_full.log 2>&1; echo "full capture rc=$?"; tail -2 $O/ncu_full.log
