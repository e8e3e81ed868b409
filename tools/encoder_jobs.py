#!/usr/bin/env python3
"""Whole-encoder jobs at BASELINE.json's sizes (encjob.JOBS), one JSON line per job: the job is sharded over the ranks this
script is launched with (python tools/encoder_jobs.py JOB... for one GPU; python -m torch.distributed.run --nproc-per-node N
... tools/encoder_jobs.py JOB... for N), bitstream and payload digests of every shard are checked against per-shard runs of
the reference encoder on the host cores, which are timed as the CPU side of the same line.  bench.py runs the same leg
(bench.encoder_job_leg) on a bounded job at every N; this is the full-size companion."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import bench
    bench.claim_stdout()
    import pcamv_loader
    pcamv = pcamv_loader.load()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for name in sys.argv[1:]:
        res = bench.encoder_job_leg(pcamv, name, rank, world, local_rank, barrier)
        if rank == 0:
            bench.emit(res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
