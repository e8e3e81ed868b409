"""Whole-encoder jobs (video-steganography-pcamv_b200/encjob.py): the CPU side — job clip generation by parallel generator
processes, per-shard reference runs with cached digests — and, on the GPU box, the sharded GPU encoder against them."""
import hashlib
import os

import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not refrun.have_ref("x264_wide"), reason="oracle/_ref/x264_wide not built")
def test_job_clip_and_reference_side(pcamv, tmp_path, monkeypatch):
    from pcamv_b200 import encjob
    monkeypatch.setenv("PCAMV_JOB_DIR", str(tmp_path))
    job = encjob.JOBS["tiny"]
    clip = encjob.make_clip(pcamv, "tiny", workers=2)
    frame = job["width"] * job["height"] * 3 // 2
    assert os.path.getsize(clip) == frame * job["shards"] * job["shard_frames"]
    # shard g of the clip is the generator's stream g, wherever it was written from
    raw = open(clip, "rb").read()
    for g in (0, job["shards"] - 1):
        alone = refrun.synth_clip(pcamv, job["width"], job["height"], job["shard_frames"], config=job["synth"], stream=g, workdir=str(tmp_path))
        k = frame * job["shard_frames"]
        assert hashlib.md5(raw[g * k:(g + 1) * k]).hexdigest() == hashlib.md5(open(alone, "rb").read()).hexdigest()
    ref = encjob.reference_side(pcamv, "tiny", cores=4)
    assert len(ref["md5"]) == job["shards"] and len(set(ref["md5"])) == job["shards"]         # different content per shard
    assert ref["frames"] == 16 and ref["fps"] > 0 and sum(ref["payload_bits"]) > 100
    again = encjob.reference_side(pcamv, "tiny")                                               # cached
    assert again["md5"] == ref["md5"] and again["seconds"] == ref["seconds"]
    assert encjob.job_args(job)[-2:] == ["--keyint", "4"]


@pytest.mark.gpu
@pytest.mark.skipif(not refrun.have_ref("x264_wide"), reason="oracle/_ref/x264_wide not built")
def test_sharded_job_matches_reference_side(pcamv, cuda_lib, tmp_path, monkeypatch):
    """One rank of a two-rank deal (shards 1 and 3) and the single-rank run: NAL stream and payload digests of every shard equal
    the per-shard reference runs'."""
    from pcamv_b200 import encjob
    monkeypatch.setenv("PCAMV_JOB_DIR", str(tmp_path))
    encjob.make_clip(pcamv, "tiny", workers=4)
    ref = encjob.reference_side(pcamv, "tiny", cores=4)
    for rank, world in ((0, 1), (1, 2)):
        _, recs, _ = encjob.run_rank("tiny", rank, world, 0)
        assert [r["gop"] for r in recs] == list(range(rank, 4, world))
        for r in recs:
            assert r["md5"] == ref["md5"][r["gop"]], "shard %d: bitstream differs" % r["gop"]
            assert r["payload_md5"] == ref["payload_md5"][r["gop"]] and r["n_bits"] == ref["payload_bits"][r["gop"]]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "x264_dump_conformant")) or not os.path.exists(os.path.join(ROOT, "host", "_build", "x264_pcamv")),
                    reason="oracle/_ref/x264_dump_conformant or host/_build/x264_pcamv not built")
def test_round_trip_job_through_the_bitstream_alone(tmp_path):
    """tools/round_trip_job.py on the CPU-sized job, with the conformant reference as the encoder: every GOP's payload read back from
    its .264 alone is the embedded message (the GPU run of the same tool encodes with x264_pcamv, PCAMV_CONFORMANT=1)."""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "round_trip_job.py"), "tiny", "--encoder", "reference"], capture_output=True, text=True,
                       env=dict(os.environ, PCAMV_JOB_DIR=str(tmp_path)), timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["payload_equals_embedded_message"] and res["p_frames_extracted"] == 12 and res["payload_bits"] > 500 and res["frames_without_payload"] == 0
