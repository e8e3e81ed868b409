#!/bin/bash
# compute-sanitizer over the hot path: memcheck / initcheck / synccheck on smoke() (border, half-pel filter,
# wavefront, cost table on a QCIF frame), racecheck (shared-memory hazards) on the multi-context launches — row groups, row pool,
# split wavefront — and on the embed stage.  Logs go to gpurun_out/san_*.log; summaries are copied to profiles/.
cd $GRAFT_REPO_ROOT
O=gpurun_out
S="compute-sanitizer --print-limit 20 --error-exitcode 9"
SMOKE='import __graft_entry__ as g; g.smoke()'
timeout 900 $S --tool memcheck  --log-file $O/san_memcheck_smoke.log  python -c "$SMOKE" > $O/san_memcheck_smoke.out 2>&1; echo "memcheck smoke rc=$?"
timeout 900 $S --tool initcheck --log-file $O/san_initcheck_smoke.log python -c "$SMOKE" > $O/san_initcheck_smoke.out 2>&1; echo "initcheck smoke rc=$?"
timeout 900 $S --tool synccheck --log-file $O/san_synccheck_smoke.log python -c "$SMOKE" > $O/san_synccheck_smoke.out 2>&1; echo "synccheck smoke rc=$?"
timeout 1500 $S --tool racecheck --log-file $O/san_racecheck_batch.log python -m pytest tests/test_gpu_frame.py -m gpu -x -q -k "batch_launch_equals_single" > $O/san_racecheck_batch.out 2>&1; echo "racecheck batch rc=$?"
timeout 900 $S --tool memcheck --log-file $O/san_memcheck_batch.log python -m pytest tests/test_gpu_frame.py -m gpu -x -q -k "batch_launch_equals_single or batch_pass2" > $O/san_memcheck_batch.out 2>&1; echo "memcheck batch rc=$?"
timeout 900 $S --tool memcheck --log-file $O/san_memcheck_embed.log python -m pytest tests/test_gpu_embed.py tests/test_gpu_stc.py -m gpu -x -q -k "golden" > $O/san_memcheck_embed.out 2>&1; echo "memcheck embed rc=$?"
timeout 900 $S --tool racecheck --log-file $O/san_racecheck_embed.log python -m pytest tests/test_gpu_embed.py -m gpu -x -q -k "golden and hex5" > $O/san_racecheck_embed.out 2>&1; echo "racecheck embed rc=$?"
for f in $O/san_*.log; do echo "== $f"; grep -c "=========" $f; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|hazard" $f | sort | uniq -c | head -8; done
tail -3 $O/san_racecheck_batch.out
