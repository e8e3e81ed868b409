#!/bin/bash
# round 2, GPU call 1: instruction-fetch probe, baseline A/B timing, search-seam timing, ncu captures of the cost-table kernel
cd $GRAFT_REPO_ROOT
O=gpurun_out
export PCAMV_QT_DIR=/tmp/pcamv_qt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/c1_smi.txt
timeout 300 build/icache_probe 105 > $O/c1_icache_probe.txt 2>&1; echo "icache rc=$?"
timeout 900 python tools/quick_time.py 128 4 2 > $O/c1_qt.log 2>&1; echo "qt rc=$?"; tail -1 $O/c1_qt.log
timeout 600 python tools/probes/search_probe.py 16 > $O/c1_search_probe.log 2>&1; echo "search probe rc=$?"; tail -2 $O/c1_search_probe.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_cost_table_batch -c 1 -f -o $O/prof_ct128_r02_base \
    python tools/quick_time.py 128 4 1 > $O/c1_ncu_ct.log 2>&1; echo "ncu ct rc=$?"; tail -2 $O/c1_ncu_ct.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_search_batch -s 3 -c 1 -f -o $O/prof_search_r02_base \
    python tools/probes/search_probe.py 16 > $O/c1_ncu_sb.log 2>&1; echo "ncu sb rc=$?"; tail -2 $O/c1_ncu_sb.log
ls -la $O
