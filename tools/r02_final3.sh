#!/bin/bash
# end-of-round validation after the last session of round 2 (decoder side, conformance switch): the whole GPU suite on 8 pytest
# workers, smoke, the conformant-mode round trips of BASELINE configs 5 and 2 through the .264 alone, both bench arms.
# (Run as:  gpurun --timeout 900 -- 'bash tools/r02_final3.sh' ; the session itself ran the pieces separately: r02_call29 .. 32.)
cd $GRAFT_REPO_ROOT
O=$PWD/gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -n 8 > $O/final3_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/final3_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $O/final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final3_smoke.log | cut -c1-300
PCAMV_JOB_DIR=/dev/shm/pcamv_jobs python tools/round_trip_job.py config5 > $O/final3_round_trip_config5.json 2> $O/final3_rt5.err; echo "config5 rc=$?"; cut -c1-600 $O/final3_round_trip_config5.json
rm -rf /dev/shm/pcamv_jobs
python bench.py > $O/final3_bench.json 2> $O/final3_bench.err; echo "bench rc=$?"; cut -c1-300 $O/final3_bench.json
python bench.py --impl reference > $O/final3_bench_reference.json 2> $O/final3_ref.err; echo "ref arm rc=$?"
