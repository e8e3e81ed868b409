#!/bin/bash
# what is left of the GPU budget: random sweep of the bound host in conformant mode against x264_dump_conformant
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
mkdir -p $O
timeout 80 python tools/host_sweep.py 23 60 conformant > $O/r02_conformant_sweep.txt 2>&1; echo "sweep rc=$?"
grep -c "^OK" $O/r02_conformant_sweep.txt; grep -c "^DIFF" $O/r02_conformant_sweep.txt; grep "^DIFF" $O/r02_conformant_sweep.txt | head -5 | cut -c1-400; tail -2 $O/r02_conformant_sweep.txt | cut -c1-300
