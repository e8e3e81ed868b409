#!/usr/bin/env python3
"""bench.py — ME Mcandidates/s (and analysed P-frames/s) of the PCAMV motion-estimation hot path on B200.

One "step" = the hot path over one 1080p P-frame exactly as the encoder needs it (BASELINE.json config 2):
  pass 1  macroblock-wavefront analysis (every x264_me_search_ref / x264_me_refine_qpel of the frame, P_SKIP probes,
          mode decision) + the PCAMV candidate-MV cost table of every motion vector (x264_ih_get_mv_cost)
  pass 2  the wavefront analysis again with the pass-1 decisions and the STC flips forced
Candidates (block-distortion evaluations at one MV) are credited as the REFERENCE executes them for the same
frame — per-pass counters of the instrumented reference (oracle/ref_hooks.c: sad/satd +1, sad_x3 +3, sad_x4 +4,
MV_SATD_FDEC_IH +1 luma +2 chroma) — never as the GPU happens to evaluate them.

  value   device-resident: frame inputs already in HBM; CUDA-event time of the three kernels on the context's stream
  e2e     through the C-ABI with HOST buffers: pcamv_put_fenc + pcamv_put_ref (H2D, GPU border + half-pel filter),
          pcamv_analyse_p pass 1 (H2D co-located MVs, D2H records + log), pcamv_analyse_p pass 2 (H2D pass-1
          records + flips, D2H records + log) — wall clock around the calls, copies inside the timed region
  --impl reference   the reference's own CPU implementation (oracle/_ref/x264_dump, built from /root/reference's C
          sources) on the box's host cores, same clip and flags: reference-counted candidates / seconds spent
          inside x264_macroblock_analyse of the P slices

A parity gate runs before any timing: records, search logs and cost table of both passes must equal the
reference's bit for bit (tests/frame_parity.py), otherwise bench.py exits non-zero.

Launch: python bench.py [--gpus N --steps K --warmup W]   (N>1 via torch.distributed.run, one rank per GPU; every
rank analyses its own clip = an independent GOP shard; no data-path collective, weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WIDTH, HEIGHT = 1920, 1080
REF_ARGS = "--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2"
WORKLOAD = ("1080p synthetic YUV420 (synth/pcamv_synth.c config 2), --me umh --subme 5 --ref 1 --qp 26 --emrate 0.2; "
            "subme 5 instead of 7: RD mode decision is raster-serial on CABAC state (DESIGN.md)")
DISTINCT = 8               # distinct P frames the contexts of a GPU analyse (context i gets frame BATCH_FRAME + i % DISTINCT of its clip)
CLIP_FRAMES = 3            # I P P: the clip of the CPU legs (reference arm, cpu_baseline)
BATCH_FRAME = 2            # first analysed frame: spatial + temporal MV candidates are both live from here on
METRIC = "me_mcandidates_per_sec"
ENCODER_JOB = "config4-small"      # the whole-encoder job every default run encodes across its ranks (encjob.JOBS)


def prepare_inputs(pcamv, rank, workdir, n_frames=1):
    """Synthetic clip + instrumented reference run for frames BATCH_FRAME .. BATCH_FRAME + n_frames - 1: planes, expected
    records, work counters."""
    import refrun
    clip = refrun.synth_clip(pcamv, WIDTH, HEIGHT, BATCH_FRAME + n_frames, config=2, stream=rank, workdir=workdir)
    dump = os.path.join(workdir, "dump.bin")
    refrun.run_ref(clip, WIDTH, HEIGHT, REF_ARGS.split(), dump=dump, frames="%d:%d" % (BATCH_FRAME, BATCH_FRAME + n_frames), count=True)
    return clip, dump


def cpu_reference_timing(pcamv, clip, workdir):
    """Clean timing run of the reference (no counting wrappers, no dump)."""
    import refrun
    stats = os.path.join(workdir, "stats_time.json")
    refrun.run_ref(clip, WIDTH, HEIGHT, REF_ARGS.split(), stats=stats)
    return json.load(open(stats))


def cpu_reference_counts(pcamv, clip, workdir):
    import refrun
    stats = os.path.join(workdir, "stats_count.json")
    refrun.run_ref(clip, WIDTH, HEIGHT, REF_ARGS.split(), stats=stats, count=True)
    return json.load(open(stats))


def candidates_of(c):
    return c["sad"] + c["satd"] + c["ih_luma"] + c["ih_chroma"]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.stop_flag, self.max_mhz = gpu, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline_block(pcamv, clip, workdir):
    """The reference's CPU path on this box: candidates of the clip's P frames / time inside x264_macroblock_analyse."""
    counts = cpu_reference_counts(pcamv, clip, workdir)
    st = cpu_reference_timing(pcamv, clip, workdir)
    cand = candidates_of(counts)
    return {"value": cand / st["t_analyse_p"] / 1e6, "unit": "Mcandidates/s", "cores": 1, "kind": "reference",
            "sample": "oracle/_ref/x264_dump (reference C sources, gcc -O4 -ffast-math, no asm; the encoder is single-threaded "
                      "with embedding on), %d-frame 1080p clip, %d candidates in %.2f s inside x264_macroblock_analyse of the "
                      "P slices (search %.2f s, cost table %.2f s); whole encode %.2f s = %.2f frames/s"
                      % (CLIP_FRAMES, cand, st["t_analyse_p"], st["t_me"], st["t_ih"], st["t_total"], CLIP_FRAMES / st["t_total"]),
            "analysed_p_frames_per_sec": counts["p_frames"] / st["t_analyse_p"],
            "encode_frames_per_sec": CLIP_FRAMES / st["t_total"]}


def _x264_fps(stderr_bytes):
    """fps values of the CLI's closing lines 'encoded N frames, X fps' (x264.c:928): the encode loop alone, without process
    start, encoder open and teardown."""
    import re
    return [float(m.group(1)) for m in re.finditer(rb"encoded \d+ frames, ([0-9.]+) fps", stderr_bytes)]


def _run_encoder(binary, args, out, clip, env=None):
    t0 = time.perf_counter()
    p = subprocess.run([binary] + list(args) + ["-o", out, clip, "%dx%d" % (WIDTH, HEIGHT)], env=env, capture_output=True)
    return time.perf_counter() - t0, _x264_fps(p.stderr), p


ESA_ARGS = "--qp 26 --ref 4 --keyint 250 --me esa --merange 32 --subme 5 --emrate 0.2"       # BASELINE.json config 3


def encoder_e2e(pcamv, workdir, device, frames=24, ref_args=REF_ARGS, config=2, tag="e2e"):
    """Whole-encoder leg: the reference's C host with the CUDA shim bound in (host/_build/x264_pcamv) against the
    reference encoder on the same 1080p clip and flags — frames/s of encode + embed, bitstreams compared."""
    import hashlib
    import refrun
    host = os.path.join(ROOT, "host", "_build", "x264_pcamv")
    ref = os.path.join(ROOT, "oracle", "_ref", "x264_wide")
    if not os.path.exists(host):
        return {"unavailable": "host/_build/x264_pcamv is not built"}
    clip = refrun.synth_clip(pcamv, WIDTH, HEIGHT, frames, config=config, stream=0, workdir=workdir)
    ref_out, out = os.path.join(workdir, tag + "_ref.264"), os.path.join(workdir, tag + "_gpu.264")
    t_ref, fps_ref, _ = _run_encoder(ref, ref_args.split(), ref_out, clip)
    stats = os.path.join(workdir, tag + "_stats.json")
    t_gpu, fps_gpu, p = _run_encoder(host, ref_args.split(), out, clip, env=dict(os.environ, PCAMV_STATS=stats, PCAMV_DEVICE=str(device)))
    if p.returncode != 0:
        return {"unavailable": "x264_pcamv failed: " + p.stderr[-300:].decode("latin-1")}
    st = json.load(open(stats))
    same = hashlib.md5(open(out, "rb").read()).hexdigest() == hashlib.md5(open(ref_out, "rb").read()).hexdigest()
    return {"frames": frames, "args": ref_args, "bitstream_identical": same,
            "reference_fps": fps_ref[0] if fps_ref else None, "ours_fps": fps_gpu[0] if fps_gpu else None,
            "reference_fps_wall": frames / t_ref, "ours_fps_wall": frames / t_gpu,
            "ours_seconds_in_gpu_calls": st["t_gpu_calls"], "ours_seconds_open": st.get("t_open"), "ours_seconds_total": st["t_total"],
            "note": "one encoder process, one stream; *_fps = the CLI's own 'encoded N frames, X fps' line (encode loop), *_fps_wall = "
                    "whole process incl. CUDA context creation; host entropy coding / reconstruction / deblocking / STC embedding are "
                    "the reference's own C code in both"}


def encoder_e2e_sharded(pcamv, workdir, device, frames_per_shard=12):
    """Whole-encoder throughput of one GPU + the box's host cores on IDR-bounded shards: `x264_pcamv --shards N` (N = two
    encoder threads per core in 4 rendezvous groups, so that host work and GPU launches of different groups overlap)
    against the reference encoder run as one process per core over the same shards (`--seek g*K --frames K`); outputs
    compared (concatenation in GOP order)."""
    import hashlib
    import refrun
    from concurrent.futures import ThreadPoolExecutor
    host = os.path.join(ROOT, "host", "_build", "x264_pcamv")
    ref = os.path.join(ROOT, "oracle", "_ref", "x264_wide")
    if not os.path.exists(host):
        return {"unavailable": "host/_build/x264_pcamv is not built"}
    cores = max(1, min(os.cpu_count() or 1, 16))
    n, k = 2 * cores, frames_per_shard
    clip = refrun.synth_clip(pcamv, WIDTH, HEIGHT, n * k, config=2, stream=0, workdir=workdir)
    ref_outs = [os.path.join(workdir, "shard_ref_%d.264" % g) for g in range(n)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda g: _run_encoder(ref, REF_ARGS.split() + ["--seek", str(g * k), "--frames", str(k)], ref_outs[g], clip), range(n)))
    t_ref = time.perf_counter() - t0
    out = os.path.join(workdir, "shard_gpu.264")
    env = dict(os.environ, PCAMV_DEVICE=str(device), PCAMV_ROWS_PER_CTA="4", PCAMV_GROUPS="4")
    env.pop("CUDA_DEVICE_MAX_CONNECTIONS", None)
    t_gpu, fps_gpu, p = _run_encoder(host, ["--shards", str(n), "--shard-frames", str(k)] + REF_ARGS.split(), out, clip, env=env)
    if p.returncode != 0:
        return {"unavailable": "x264_pcamv --shards failed: " + p.stderr[-300:].decode("latin-1")}
    want = hashlib.md5(b"".join(open(f, "rb").read() for f in ref_outs)).hexdigest()
    same = hashlib.md5(open(out, "rb").read()).hexdigest() == want
    steady = n * k / max(k / f for f in fps_gpu) if len(fps_gpu) == n else None
    return {"shards": n, "frames_per_shard": k, "frames": n * k, "host_cores": cores, "bitstream_identical": same,
            "reference_fps": n * k / t_ref, "ours_fps": steady, "ours_fps_wall": n * k / t_gpu,
            "note": "reference_fps = frames / wall clock of %d concurrent single-threaded reference encoder processes (one per host core) "
                    "working through the shards; ours_fps = frames / the slowest shard's encode-loop time (the CLI's own 'encoded N "
                    "frames, X fps' lines; all shards run concurrently in one process on one GPU), ours_fps_wall = frames / wall clock "
                    "of the whole job incl. process start, CUDA context creation and teardown, which %d-frame shards do not amortise"
                    % (cores, k)}


def run_reference_arm(args, pcamv, rank, world):
    """The reference's own CPU implementation with all the host cores it can use: the encoder is single-threaded with
    embedding on (frame threads crash), so one process per core, each encoding its own clip (an
    independent GOP shard / stream), all running concurrently."""
    if rank != 0:
        return
    import refrun
    from concurrent.futures import ThreadPoolExecutor
    cores = max(1, min(os.cpu_count() or 1, 64))
    workdir = tempfile.mkdtemp(prefix="pcamv_bench_ref_")
    pcamv.build.build_synth()            # once, before the worker threads need it
    dirs = []
    for i in range(cores):
        d = os.path.join(workdir, "s%d" % i)
        os.makedirs(d)
        dirs.append(d)
    with ThreadPoolExecutor(cores) as ex:
        clips = list(ex.map(lambda i: refrun.synth_clip(pcamv, WIDTH, HEIGHT, CLIP_FRAMES, config=2, stream=i, workdir=dirs[i]), range(cores)))
        counts = list(ex.map(lambda i: cpu_reference_counts(pcamv, clips[i], dirs[i]), range(cores)))
        cand = sum(candidates_of(c) for c in counts)
        p_frames = sum(c["p_frames"] for c in counts)
        times, totals = [], []
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            st = list(ex.map(lambda i: cpu_reference_timing(pcamv, clips[i], dirs[i]), range(cores)))
            wall = time.perf_counter() - t0
            if k >= args.warmup:
                times.append(max(s["t_analyse_p"] for s in st))      # slowest core's time inside the analysis
                totals.append(wall)
    t = float(np.mean(times))
    v = cand / t / 1e6
    sample = ("%d concurrent reference encoder processes (one per host core), each a whole %d-frame 1080p clip per step (%d P frames, both "
              "passes each, %d reference-counted candidates in total); time = the slowest process's seconds inside x264_macroblock_analyse "
              "of the P slices (its border expansion and half-pel filter, x264_frame_filter, are NOT in this time although the GPU arm's "
              "e2e includes them through pcamv_put_ref: the time base favours the reference)" % (cores, CLIP_FRAMES, p_frames, cand))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mcandidates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 / (p_frames / cores), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames": CLIP_FRAMES, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "Mcandidates/s", "cores": cores, "kind": "reference",
                             "sample": "oracle/_ref/x264_dump (reference C sources, gcc -O4 -ffast-math, no asm; single encoder thread per "
                                       "process: frame threads crash with embedding on), " + sample},
            "analysed_p_frames_per_sec": p_frames / t,
            "encode_frames_per_sec": cores * CLIP_FRAMES / float(np.mean(totals)),
            "e2e": {"value": v, "unit": "Mcandidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.encoder_job != "none":
        # the same whole-encoder job the GPU arm shards over its ranks, here as one reference encoder process per shard on
        # the host cores; the outputs' digests stay cached on this box as the parity oracle of the GPU arm
        from pcamv_b200 import encjob
        encjob.make_clip(pcamv, args.encoder_job)
        ref = encjob.reference_side(pcamv, args.encoder_job)
        job = encjob.JOBS[args.encoder_job]
        line["encoder_job"] = {"job": args.encoder_job, "what": job["what"], "args": " ".join(encjob.job_args(job)), "frames": ref["frames"],
                               "shards": job["shards"], "encode_embed_fps": ref["fps"], "seconds": ref["seconds"], "cores": ref["cores"],
                               "note": "oracle/_ref/x264_wide (the reference's C sources, no asm), one single-threaded process per shard"}
    emit(line)


_RESULT_OUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the result.  Everything else that writes to file descriptor 1 — NCCL's version banner
    (NCCL_DEBUG from the environment or from /etc/nccl.conf), child processes, libraries — is sent to stderr: fd 1 is
    duplicated for the result line and then pointed at fd 2."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--streams", type=int, default=128,
                    help="independent encoder contexts (GOP shards / streams) analysing concurrently on each GPU")
    ap.add_argument("--launch", choices=["batch", "streams"], default="batch",
                    help="batch: all contexts' frames in ONE wavefront launch (pcamv_analyse_p_batch); streams: one launch per context on its own CUDA stream")
    ap.add_argument("--e2e-lanes", type=int, default=2, help="end-to-end arm: independent context sets whose copies and launches overlap")
    ap.add_argument("--rows-per-cta", type=int, default=4, help="wavefront layout used when --streams > 1 (pcamv_cfg.rows_per_cta)")
    ap.add_argument("--pass2-elide", action="store_true",
                    help="pcamv_cfg.pass2_elide: skip the pass-2 searches whose results the reference overwrites (not the default: "
                         "the headline number executes everything the reference executes)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference leg (profiling runs)")
    ap.add_argument("--distinct-frames", type=int, default=DISTINCT, help="distinct P frames of the clip spread over the contexts of a GPU")
    ap.add_argument("--encoder-job", default=ENCODER_JOB,
                    help="whole-encoder job sharded over the ranks (encjob.JOBS: config2 / config4 / config4-k30 / config5 at BASELINE's sizes, "
                         "*-small bounded versions), or 'none'")
    args = ap.parse_args()

    # one hardware work queue per context stream (the default of 8 would serialise contexts 9..S behind the first 8)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

    import pcamv_loader
    pcamv = pcamv_loader.load()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, pcamv, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION / INFO in the environment) goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import frame_parity
    S = max(1, args.streams)
    D = max(1, min(args.distinct_frames, S))
    workdir = tempfile.mkdtemp(prefix="pcamv_bench_%d_" % rank)
    clip, dumpf = prepare_inputs(pcamv, rank, workdir, n_frames=D)
    dump = pcamv.dumpfmt.Dump(dumpf)
    all_units = [u for u in dump.slice_units() if u["slice"].with_planes]
    frames = sorted({u["slice"].frame for u in all_units})
    assert frames == list(range(BATCH_FRAME, BATCH_FRAME + D)), "dump holds frames %s" % frames
    by_frame = []
    for fr in frames:
        us = [u for u in all_units if u["slice"].frame == fr]
        assert [u["slice"].pass_ for u in us] == [1, 2], "dump does not hold both passes of frame %d" % fr
        by_frame.append(us)
    s = by_frame[0][0]["slice"]

    # reference-counted work per distinct frame: the per-pass counters of its two dumped passes (file order = frame order)
    cnt = dump.counters()
    assert len(cnt) == 2 * D, "expected the counters of %d passes, got %d" % (2 * D, len(cnt))
    W_SAD, W_SATD, W_AVG, W_CHROMA, W_DCT = 2.0, 7.0, 2.0, 8.0, 14.0       # integer ops per pixel, SURVEY.md 8(d)
    def ops_of(c, part):
        """reference-counted integer ops of one pass: part = 'search' (wavefront kernel) or 'table' (cost-table kernel)"""
        if part == "table":
            return W_SATD * c["ih_pix_satd"] + W_AVG * c["ih_pix_avg"] + W_CHROMA * c["ih_pix_chroma_mc"] + W_DCT * c["ih_pix_dct"]
        return (W_SAD * c["pix_sad"] + W_SATD * (c["pix_satd"] - c["ih_pix_satd"]) + W_AVG * (c["pix_avg"] - c["ih_pix_avg"]) +
                W_CHROMA * (c["pix_chroma_mc"] - c["ih_pix_chroma_mc"]) + W_DCT * (c["pix_dct"] - c["ih_pix_dct"]))
    cand_frame = [candidates_of({k: cnt[2 * d][k] + cnt[2 * d + 1][k] for k in ("sad", "satd", "ih_luma", "ih_chroma")}) for d in range(D)]
    ops_frame = [{"pass1": ops_of(cnt[2 * d], "search"), "table": ops_of(cnt[2 * d], "table"), "pass2": ops_of(cnt[2 * d + 1], "search")}
                 for d in range(D)]
    frame_of = [i % D for i in range(S)]                        # context i analyses distinct frame i % D
    cand_step = float(sum(cand_frame[d] for d in frame_of))      # candidates of one step of this rank
    ops_step = {k: float(sum(ops_frame[d][k] for d in frame_of)) for k in ("pass1", "table", "pass2")}

    # ---- parity gate before any number: every distinct frame, both passes, bit for bit against the reference's records ----
    ctxs = [frame_parity.open_ctx(pcamv, dump, s, device=local_rank, rows_per_cta=args.rows_per_cta if S > 1 else 1, pass2_elide=int(args.pass2_elide)) for _ in range(S)]
    gate = frame_parity.open_ctx(pcamv, dump, s, device=local_rank)           # (executes everything, whatever --pass2-elide says)
    par = {"calls": 0, "mbs": 0, "ih": 0}
    for us in by_frame:
        p_ = frame_parity.check_dump(pcamv, dump, units=us, ctx=gate, keep_ctx=True)       # raises on the first mismatch
        for k in par:
            par[k] += p_[k]
    gate.close()

    H, W = s.lines_y, s.width
    n_mb = (W // 16) * (H // 16)
    inputs = []
    for us in by_frame:
        s1, x, e = us[0]["slice"], us[0]["ctx"], us[0]["embd"]
        r = s1.refs[0]
        planes = [np.ascontiguousarray(s1.fenc[0][:, :W]), np.ascontiguousarray(s1.fenc[1][:, :W // 2]), np.ascontiguousarray(s1.fenc[2][:, :W // 2]),
                  np.ascontiguousarray(r["luma"][0][32:32 + H, 32:32 + W]), np.ascontiguousarray(r["u"][16:16 + H // 2, 16:16 + W // 2]),
                  np.ascontiguousarray(r["v"][16:16 + H // 2, 16:16 + W // 2])]
        # the end-to-end arm copies from / to page-locked host memory (pcamv_host_alloc), as an encoder host would hold its frames
        planes = [pcamv.host.pinned_copy(a) for a in planes]
        inputs.append(dict(fenc=planes[:3], ref=planes[3:], poc=r["poc"],
                           col=dict(col_n_ref=x["col_n_ref"], col_inv_ref_poc=x["col_inv_ref_poc"], col_ref8=x["col_ref8"], col_mv4=x["col_mv4"]),
                           pass1=frame_parity.pass1_records(pcamv, e), filp=e["filp"], refs=list(range(x["n_ref"])),
                           pocs=x["ref_poc"][:x["n_ref"]], cur_poc=x["cur_poc"]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_threads(fn):
        """fn(i) on one host thread per context (the C-ABI calls release the GIL), all started together."""
        out = [None] * S
        def wrap(i):
            out[i] = fn(i)
        th = [threading.Thread(target=wrap, args=(i,)) for i in range(S)]
        for t_ in th: t_.start()
        for t_ in th: t_.join()
        return out

    # ---- device-resident arm: every context holds its own frame inputs in HBM ----------------------------------------------
    def stage(i):
        c, q = ctxs[i], inputs[frame_of[i]]
        c.put_fenc(*q["fenc"])
        c.put_ref(0, q["poc"], *q["ref"])
        c.frame_upload(1, q["refs"], q["pocs"], q["cur_poc"], cost_table=True, **q["col"])
        m1, _ = c.analyse_p(1, q["refs"], q["pocs"], q["cur_poc"], cost_table=True, **q["col"])
        c.frame_upload(2, q["refs"], q["pocs"], q["cur_poc"], pass1=q["pass1"], filp=q["filp"], stale_mv=m1["mv"][-1], **q["col"])
        return m1
    staged = run_threads(stage)
    assert all((staged[i]["mv"] == staged[frame_of[i]]["mv"]).all() for i in range(S))        # same frame, same result
    assert D == 1 or not (staged[0]["mv"] == staged[1]["mv"]).all()                            # the frames really differ

    def dev_steps(i, k):
        c = ctxs[i]
        acc = np.zeros(3)
        for _ in range(k):
            _, w1, ct = c.frame_run(1, 1, per_kernel=True)      # CUDA events around each kernel on the context's stream
            _, w2, _ = c.frame_run(2, 1, per_kernel=True)
            acc += (w1, ct, w2)
        return acc
    def dev_steps_batch(k):
        acc = np.zeros(3)
        for _ in range(k):
            _, w1, ct = pcamv.host.frame_run_batch(ctxs, 1)      # CUDA events around each kernel on the leader's stream
            _, w2, _ = pcamv.host.frame_run_batch(ctxs, 2)
            acc += (w1, ct, w2)
        return [acc]
    batch = args.launch == "batch" and S > 1
    dev_run = (lambda k: dev_steps_batch(k)) if batch else (lambda k: run_threads(lambda i: dev_steps(i, k)))
    dev_run(args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = sum(c.launch_count() for c in ctxs)
    t0 = time.perf_counter()
    k_ms = dev_run(args.steps)
    torch.cuda.synchronize()
    dev_s = time.perf_counter() - t0
    barrier()
    k_ms = np.array(k_ms) / args.steps            # mean kernel durations (per context while all S run concurrently, or of the batch launch)
    n_launch = sum(c.launch_count() for c in ctxs) - launches0

    # ---- end-to-end arm: host buffers in, host records out ---------------------------------------------------------------------
    outs1 = [c.alloc_outputs(pinned=True) for c in ctxs]
    outs2 = [c.alloc_outputs(pinned=True) for c in ctxs]

    def e2e_steps(i, k):
        c, q = ctxs[i], inputs[frame_of[i]]
        for _ in range(k):
            c.put_fenc(*q["fenc"])
            c.put_ref(0, q["poc"], *q["ref"])
            m1, l1 = c.analyse_p(1, q["refs"], q["pocs"], q["cur_poc"], cost_table=True, out=outs1[i], **q["col"])
            m2, l2 = c.analyse_p(2, q["refs"], q["pocs"], q["cur_poc"], pass1=q["pass1"], filp=q["filp"], stale_mv=m1["mv"][-1], out=outs2[i], **q["col"])
        return m1, l1, m2, l2
    # batch launches: the contexts are split into E2E_LANES independent sets, each driven by its own host thread through
    # upload -> pcamv_analyse_p_batch(pass 1) -> pcamv_analyse_p_batch(pass 2) -> download, so that one set's PCIe copies
    # overlap the other sets' kernels
    lanes = max(1, min(args.e2e_lanes, S))
    lane_sets = [list(range(i, S, lanes)) for i in range(lanes)]

    def e2e_lane(ids, k):
        cs = [ctxs[i] for i in ids]
        qs = [inputs[frame_of[i]] for i in ids]
        for _ in range(k):
            for c, q in zip(cs, qs):
                c.put_fenc(*q["fenc"])
                c.put_ref(0, q["poc"], *q["ref"])
            a1 = [(1, q["refs"], q["pocs"], q["cur_poc"], dict(cost_table=True, **q["col"])) for q in qs]
            o1 = pcamv.host.analyse_p_batch(cs, a1, outs=[outs1[i] for i in ids])
            a2 = [(2, q["refs"], q["pocs"], q["cur_poc"], dict(pass1=q["pass1"], filp=q["filp"], stale_mv=o[0]["mv"][-1], **q["col"])) for q, o in zip(qs, o1)]
            o2 = pcamv.host.analyse_p_batch(cs, a2, outs=[outs2[i] for i in ids])
        return o1, o2

    def e2e_steps_batch(k):
        res = [None] * lanes
        def wrap(j):
            res[j] = e2e_lane(lane_sets[j], k)
        th = [threading.Thread(target=wrap, args=(j,)) for j in range(lanes)]
        for x_ in th: x_.start()
        for x_ in th: x_.join()
        out = [None] * S
        for j, ids in enumerate(lane_sets):
            o1, o2 = res[j]
            for n_, i in enumerate(ids):
                out[i] = (o1[n_][0], o1[n_][1], o2[n_][0], o2[n_][1])
        return out
    e2e_run = e2e_steps_batch if batch else (lambda k: run_threads(lambda i: e2e_steps(i, k)))
    e2e_run(args.warmup)
    barrier()
    t0 = time.perf_counter()
    outs = e2e_run(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.result()
    n_launch_e2e = sum(c.launch_count() for c in ctxs) - launches0 - n_launch
    assert all((outs[i][2]["mv"] == outs[frame_of[i]][2]["mv"]).all() and (outs[i][2]["type"] == outs[frame_of[i]][2]["type"]).all() for i in range(S))
    q0 = inputs[0]
    h2d = 2 * sum(a.nbytes for a in q0["fenc"]) + sum(a.nbytes for a in q0["ref"]) - sum(a.nbytes for a in q0["fenc"]) + 2 * 68 * n_mb + \
        n_mb * pcamv.host.PASS1_MB_DTYPE.itemsize + len(q0["filp"])
    d2h = sum(o.nbytes for o in outs[0])

    # extra, not the headline: the same device-resident step with the dead pass-2 searches elided (what the host encoder runs)
    elided = None
    if not args.pass2_elide:
        for c in ctxs:
            c.set_pass2_elide(True)
        dev_run(1)
        barrier()
        t0 = time.perf_counter()
        k2 = np.array(dev_run(3)) / 3
        torch.cuda.synchronize()
        el_s = time.perf_counter() - t0
        barrier()
        for c in ctxs:
            c.set_pass2_elide(False)
        elided = {"ms_per_step": el_s / 3 * 1e3, "steps": 3, "kernel_ms_pass2": float(k2.mean(axis=0)[2]),
                  "note": "pcamv_cfg.pass2_elide = 1: pass 2 runs only the 16x16 search of macroblocks whose decision is forced from "
                          "pass 1; the other searches' results are overwritten by the reference (encoder/analyse.c:2868-2991) and the "
                          "host encoder does not need them.  Candidates are still credited as the reference executes them."}

    int_peak = ctxs[0].int_peak_gops()
    plane_y = ctxs[0].plane_bytes(0); plane_c = ctxs[0].plane_bytes(4)
    log_stride = ctxs[0].log_stride
    for c in ctxs:
        c.close()
    ctxs = []

    # ---- whole-encoder job, sharded over the ranks: encode + embed, NCCL gather of payload / statistics inside the clock -------
    job = None
    if args.encoder_job != "none":
        job = encoder_job_leg(pcamv, args.encoder_job, rank, world, local_rank, barrier)

    # ---- aggregate over ranks (max time, summed work) ----------------------------------------------------------------
    from pcamv_b200 import shard
    dev_s_max, e2e_s_max = shard.max_over_ranks([dev_s, e2e_s])
    if elided is not None:
        elided["ms_per_step"] = shard.max_over_ranks([elided["ms_per_step"]])[0]
    cand_all = shard.sum_over_ranks([cand_step])[0]           # candidates of one step over all ranks and contexts

    if rank == 0:
        value = cand_all * args.steps / dev_s_max / 1e6
        e2e_v = cand_all * args.steps / e2e_s_max / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs_sustained", peaks.get("hbm_gbs", 6650.0))
        ms_w1, ms_ct, ms_w2 = [float(v) for v in k_ms.mean(axis=0)]
        kernels = {"k_analyse_p(pass1)": ms_w1, "k_cost_table": ms_ct, "k_analyse_p(pass2)": ms_w2}
        kops = {"k_analyse_p(pass1)": ops_step["pass1"], "k_cost_table": ops_step["table"], "k_analyse_p(pass2)": ops_step["pass2"]}
        dom = max(kernels, key=kernels.get)
        # algorithmic HBM bytes of one launch of either kernel: fenc once, 4 luma + 2 chroma reference planes once,
        # per-MB records in/out (DESIGN.md "Data layout"); a batch launch covers S frames
        alg_bytes = sum(a.nbytes for a in q0["fenc"]) + 4 * plane_y + 2 * plane_c + n_mb * (128 + log_stride * 16 + 68)
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if tr["contexts_per_gpu"] == S and batch:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]       # per launch of the dominant kernel, one ncu capture
                traffic_src = tr.get("source")
        except Exception:
            pass
        # Integer-issue roofline (this path is integer-ALU / issue bound, not HBM bound).  Numerator: the
        # REFERENCE's work for the frames of one launch in 8(d)'s per-pixel weights (SAD 2, SATD 7, quarter-pel average 2,
        # chroma bilinear 8, DCT+quant+dequant+IDCT 14), counted by the instrumented reference (oracle/ref_hooks.c CNT0 / CNT1).
        # Denominator: 32-bit lane-instructions per second of the VABSDIFF4 / IADD3 / LOP3 mix measured on this GPU
        # (pcamv_int_peak).  A packed VABSDIFF4-accumulate retires 8 numerator ops per lane-instruction, so 100 % is not the
        # ceiling of a perfectly packed SAD loop; the figure is comparable between kernels and rounds, which is its use.
        def roof(name):
            ach = kops[name] / (kernels[name] * 1e-3) / 1e9
            return {"kernel": name, "bound": "int_issue", "achieved": ach, "peak": int_peak, "unit": "Gop/s", "frac": ach / int_peak,
                    "ms_per_launch": kernels[name], "ops_per_launch": kops[name],
                    "hbm_algorithmic_gbs": S * alg_bytes / (kernels[name] * 1e-3) / 1e9}
        per_kernel = [roof(k) for k in kernels]
        dom_roof = dict(roof(dom))
        dom_roof.update({"traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": int(S * alg_bytes),
                         "hbm_peak_gbs": hbm_peak, "hbm_frac": S * alg_bytes / (kernels[dom] * 1e-3) / 1e9 / hbm_peak,
                         "peak_source": "pcamv_int_peak microbenchmark on this GPU (lane-instructions/s); HBM peak from " +
                                        ("MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"),
                         "ops": "reference-counted per launch: 2/px SAD, 7/px SATD, 2/px qpel average, 8/px chroma MC, 14/px DCT+quant+IDCT",
                         "note": "integer-issue bound (SURVEY.md 8(d)); HBM traffic is a fraction of a percent of peak by design"})
        frames_all = world * S * args.steps
        ops_all = sum(ops_step.values())
        line = {
            "metric": METRIC, "value": value, "unit": "Mcandidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_s_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "step": "one 1080p P frame through the frame seam in each of %d independent encoder contexts per GPU (GOP shards: own "
                               "frame buffers; %d distinct frames of the clip, context i holds frame %d + i %% %d): wavefront analysis pass 1 + "
                               "candidate-MV cost table + wavefront analysis pass 2" % (S, D, BATCH_FRAME, D),
                       "contexts_per_gpu": S, "distinct_frames": D,
                       "pass2": ("searches whose results the reference overwrites are elided (pcamv_cfg.pass2_elide)"
                                 if args.pass2_elide else "everything the reference executes"),
                       "launch": ("one wavefront launch for all contexts (pcamv_analyse_p_batch)" if batch else
                                  "one launch per context, each on its own CUDA stream"),
                       "candidates_per_step_per_gpu": int(cand_step),
                       "l2": "inputs larger than L2: %d contexts x %.1f MB of planes each, no flush" % (S, (alg_bytes) / 1e6),
                       "parity_gate": "passed: %d searches, %d macroblock decisions, %d cost-table entries of %d distinct frames bit-exact vs reference"
                                      % (par["calls"], par["mbs"], par["ih"], D)},
            "analysed_p_frames_per_sec": frames_all / dev_s_max,
            "kernel_ms": kernels,
            "e2e": {"value": e2e_v, "unit": "Mcandidates/s", "h2d_bytes_per_step": int(h2d) * S, "d2h_bytes_per_step": int(d2h) * S,
                    "ms_per_step": e2e_s_max / args.steps * 1e3, "analysed_p_frames_per_sec": frames_all / e2e_s_max,
                    "note": "includes pcamv_put_ref (H2D + GPU border expansion + half-pel filter), which the reference arm's clock "
                            "(x264_macroblock_analyse of the P slices) does not contain"},
            "gpu_launches": int(n_launch + n_launch_e2e),
            "clocks": clocks,
            "roofline": dom_roof,
            "roofline_per_kernel": per_kernel,
            "int_issue": {"achieved_gops": ops_all * args.steps / dev_s_max / 1e9, "peak_gops": int_peak,
                          "frac": ops_all * args.steps / dev_s_max / 1e9 / int_peak, "ops": "whole step (three kernels), per GPU, weights as in roofline"},
        }
        if elided is not None:
            elided["value"] = cand_all / (elided["ms_per_step"] * 1e-3) / 1e6
            elided["unit"] = "Mcandidates/s"
            line["pass2_elided"] = elided
        if job is not None:
            line["encoder_job"] = job
        if not args.no_cpu_baseline and world == 1:          # (CPU and single-stream legs: rank 0 at N = 1 only)
            cclip = __import__("refrun").synth_clip(pcamv, WIDTH, HEIGHT, CLIP_FRAMES, config=2, stream=0, workdir=workdir)
            line["cpu_baseline"] = cpu_baseline_block(pcamv, cclip, workdir)
            line["encoder_e2e"] = encoder_e2e(pcamv, workdir, local_rank)
            # BASELINE.json config 3 (exhaustive search, merange 32, 4 references): one stream, whole encoder
            line["encoder_e2e_esa"] = encoder_e2e(pcamv, workdir, local_rank, frames=6, ref_args=ESA_ARGS, config=3, tag="esa")
            line["round_trip"] = round_trip_leg(local_rank)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def round_trip_leg(device, job="config5-small"):
    """BASELINE config 5's wording on a bounded job - embed + extract round trip - with nothing but the .264 on the extraction
    side: tools/round_trip_job.py encodes the job's streams with x264_pcamv in conformant mode, reads every stream
    back with `x264_pcamv --extract-264` and compares the payload with the embedded message.  A side leg (its own process, bounded
    by a timeout); a failure is reported in the line, it never takes the bench down."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "round_trip_job.py"), job], capture_output=True, timeout=240,
                           env=dict(os.environ, PCAMV_DEVICE=str(device)))
        last = [ln for ln in p.stdout.decode("latin-1").splitlines() if ln.startswith("{")]
        if not last:
            return {"job": job, "error": ("rc %d: " % p.returncode) + p.stderr.decode("latin-1")[-300:]}
        return json.loads(last[-1])
    except Exception as e:                                     # noqa: BLE001 - a side leg
        return {"job": job, "error": str(e)[:300]}


def encoder_job_leg(pcamv, name, rank, world, local_rank, barrier):
    """BASELINE configs 2 / 4 / 5 through the whole encoder, sharded over the ranks (encjob.py).  Timed region: every rank's
    x264_pcamv process over its shards + the NCCL gather of per-shard payload bits, statistics and digests; max over ranks."""
    import torch
    from pcamv_b200 import encjob, shard
    job = encjob.JOBS[name]
    ref = None
    if rank == 0:
        encjob.make_clip(pcamv, name)
        ref = encjob.reference_side(pcamv, name)         # cached per box (the reference arm of bench.py fills the same cache)
    barrier()
    t0 = time.perf_counter()
    secs, recs, _ = encjob.run_rank(name, rank, world, local_rank)
    t_enc = time.perf_counter() - t0
    gathered = shard.gather_gop_results(recs)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    barrier()
    # inside the encoder processes (PCAMV_STATS of every shard): the slowest shard's lifetime without / with CUDA start-up
    st = [r["stats"] for r in recs if r.get("stats")]
    loop = max([s_["t_total"] - s_["t_open"] for s_ in st] or [0.0])
    t_open = max([s_["t_open"] for s_ in st] or [0.0])
    gpu_share = max([s_["t_gpu_calls"] / max(s_["t_total"] - s_["t_open"], 1e-9) for s_ in st] or [0.0])
    t_max, t_enc_max, loop_max, open_max, gpu_share_max = shard.max_over_ranks([t_all, t_enc, loop, t_open, gpu_share])
    # where an encoder thread of this rank spends its time, mean over its shards, per P frame (ms): pass 1 on the GPU incl. the
    # wait for the group, the embed stage, waiting for rows of the replayed pass, everything else (the host's own work)
    per_frame = None
    if st:
        nf = max(sum(s_.get("direct_pass1", 0) for s_ in st), 1)
        tot = sum(s_["t_total"] - s_["t_open"] for s_ in st)
        p1, em, rw, gp = (sum(s_.get(k, 0.0) for s_ in st) for k in ("t_pass1", "t_embed", "t_row_wait", "t_gpu_calls"))
        per_frame = {"pass1_ms": 1e3 * p1 / nf, "embed_ms": 1e3 * em / nf, "row_wait_ms": 1e3 * rw / nf,
                     "other_gpu_calls_ms": 1e3 * (gp - p1 - em - rw) / nf, "host_ms": 1e3 * (tot - gp) / nf, "frames": nf}
    if rank != 0:
        return None
    n = job["shards"]
    ok_bits = len(gathered) == n and all(g["md5"] == ref["md5"][g["gop"]] for g in gathered)
    ok_pay = None
    if ref.get("payload_md5"):
        ok_pay = len(gathered) == n and all(g["payload_md5"] == ref["payload_md5"][g["gop"]] and g["n_bits"] == ref["payload_bits"][g["gop"]] for g in gathered)
    frames = n * job["shard_frames"]
    return {"job": name, "what": job["what"], "args": " ".join(encjob.job_args(job)), "frames": frames, "shards": n, "ranks": world,
            "encode_embed_fps": frames / t_max, "seconds": t_max, "seconds_encode_only": t_enc_max, "seconds_gather": t_max - t_enc_max,
            "encode_loop_fps": frames / loop_max if loop_max > 0 else None, "seconds_encode_loop": loop_max, "seconds_cuda_startup": open_max,
            "gpu_call_share_of_encoder_thread": gpu_share_max, "encoder_thread_ms_per_p_frame_rank0": per_frame,
            "bitstream_identical": bool(ok_bits), "payload_identical": ok_pay,
            "payload_bits": int(sum(g["n_bits"] for g in gathered)), "bytes": int(sum(g["bytes"] for g in gathered)),
            "reference_fps": ref["fps"], "reference_cores": ref["cores"], "reference_seconds": ref["seconds"],
            "reference_source": ref.get("source", "per-shard runs of oracle/_ref/x264_wide on this box's host cores"),
            "scaling": "strong (the job is fixed, shards are dealt round-robin to the ranks)",
            "gather": "shard.gather_gop_results (NCCL all_gather x3) of payload bits + statistics + digests, inside the timed region" if world > 1
                      else "single rank: no collective",
            "note": "wall clock of the slowest rank's x264_pcamv process (process start, CUDA context creation and encoder open included) plus the "
                    "gather; encode_loop_fps = the same job without the slowest shard's pcamv_open (CUDA initialisation, 0.8 - 2.8 s on these boxes); reference = oracle/_ref/x264_wide (the reference's C sources, no asm), one single-threaded process per shard on the host cores"}


if __name__ == "__main__":
    main()
