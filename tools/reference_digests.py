#!/usr/bin/env python3
"""Per-shard digests of the reference encoder's output for the whole-encoder jobs (encjob.JOBS): NAL-stream md5, payload md5
and payload bits of every shard, from per-shard runs of oracle/_ref/x264_wide / x264_dump on THIS machine's cores.  The job
clips are synthetic and seeded, the reference is deterministic: the digests are the same on every box, so a multi-GPU scaling
run (charged per GPU-minute) can check parity against the committed file (PCAMV_JOB_DIGESTS=profiles/r02_reference_digests)
instead of re-running the reference there; its `reference_fps` then names the machine the file was made on.
usage: tools/reference_digests.py JOB..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    import pcamv_loader
    pcamv = pcamv_loader.load()
    from pcamv_b200 import encjob
    for name in sys.argv[1:]:
        res = encjob.reference_side(pcamv, name)
        res["made_on"] = "%d cores, %s" % (os.cpu_count(), os.uname().nodename)
        out = os.path.join(ROOT, "profiles", "r02_reference_digests", name + ".json")
        json.dump(res, open(out, "w"), indent=1)
        print(name, res["fps"], "fps on", res["cores"], "cores ->", out, flush=True)
