#!/bin/bash
# round 2, GPU call 27 (1 GPU): random option / size sweep of the bound host in its end-of-round default flow (pass 1 on the device,
# q1-only intra analysis, rows behind the wavefront, device-built references + half-pel planes), bitstream md5 against the reference
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
timeout 330 python tools/host_sweep.py 11 80 > $O/c27_sweep.txt 2>&1; echo "sweep rc=$?"
grep -c "^OK" $O/c27_sweep.txt; grep -c "^DIFF" $O/c27_sweep.txt; grep "^DIFF" $O/c27_sweep.txt | head -5 | cut -c1-400; tail -3 $O/c27_sweep.txt | cut -c1-300
