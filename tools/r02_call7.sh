#!/bin/bash
# round 2, GPU call 7: whole GPU test suite with the device-resident embed stage cross-checked against the host's vectors
cd $GRAFT_REPO_ROOT
O=gpurun_out
PCAMV_CHECK_EMBED=1 timeout 2400 python -m pytest tests -m gpu -x -q > $O/c7_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/c7_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/c7_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/c7_smoke.log
