#!/bin/bash
# last GPU call of round 2 (6.5 GPU-minutes left): the default path after the conformance-switch edits, then the new mode
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
mkdir -p $O
T0=$(date +%s)
timeout 150 python -m pytest tests/test_gpu_host.py -q -x -k "test_bitstream_identical and (cif_umh5_ref3 or cif_qp48_skips or cif_p4x4_umh_ref3 or cif_dia2_lownoise or cif_nocabac) and not host_pass1 and not full_pass2 and not switches" > $O/c29_default.log 2>&1; echo "default rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 $O/c29_default.log | cut -c1-200
timeout 150 python -m pytest tests/test_z_gpu_conformant.py -q > $O/c29_conformant.log 2>&1; echo "conformant rc=$? t=$(( $(date +%s) - T0 ))"; tail -12 $O/c29_conformant.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/c29_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))"; tail -1 $O/c29_smoke.log | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_embed.py tests/test_gpu_recon.py tests/test_gpu_frame.py -q -x > $O/c29_more.log 2>&1; echo "more rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 $O/c29_more.log | cut -c1-200
