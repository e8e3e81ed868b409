#!/bin/bash
# round 2, GPU call 14: reference frames built on the device — ABI-level parity against dumped planes, then the bound host
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_recon.py -m gpu -q > $O/c14_tests.log 2>&1; echo "tests rc=$?"; tail -25 $O/c14_tests.log | cut -c1-300
