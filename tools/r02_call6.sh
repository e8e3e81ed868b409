#!/bin/bash
# round 2, GPU call 6: cost-table variants (CTAs per SM, rolled transform), then the restructured bench (both arms, short)
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt
tools/ab.sh "default 128 4" "lib_ctp4.so 128 4" "lib_ctp6.so 128 4" "lib_ctroll.so 128 4" > $O/c6_ab.log 2>&1; cat $O/c6_ab.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/c6_bench_ref.json 2> $O/c6_bench_ref.err; echo "ref arm rc=$?"; tail -c 400 $O/c6_bench_ref.err
timeout 1500 python bench.py --steps 4 --warmup 3 > $O/c6_bench.json 2> $O/c6_bench.err; echo "bench rc=$?"; tail -c 1500 $O/c6_bench.err; head -c 3000 $O/c6_bench.json
