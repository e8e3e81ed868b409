#!/usr/bin/env python3
"""x264_pcamv --shards timing probe: python tools/shard_probe.py N K GROUPS"""
import os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n, k, g = [int(a) for a in sys.argv[1:4]]
clip = "/tmp/cs_%d.yuv" % (n * k)
if not os.path.exists(clip):
    subprocess.check_call([os.path.join(ROOT, "build", "pcamv_synth"), "1920", "1080", str(n * k), "2", "0", clip, "32"])
args = "--qp 26 --ref 1 --keyint 250 --me umh --subme 5 --emrate 0.2".split()
env = dict(os.environ, PCAMV_VERBOSE="1", PCAMV_ROWS_PER_CTA="4", PCAMV_GROUPS=str(g))
t0 = time.perf_counter()
p = subprocess.run([os.path.join(ROOT, "host", "_build", "x264_pcamv"), "--shards", str(n), "--shard-frames", str(k)] + args +
                   ["-o", "/tmp/gs.264", clip, "1920x1080"], env=env, capture_output=True, timeout=300)
wall = time.perf_counter() - t0
fps = [float(m.group(1)) for m in re.finditer(rb"encoded \d+ frames, ([0-9.]+) fps", p.stderr)]
life = [float(m.group(1)) for m in re.finditer(rb"encoder lifetime ([0-9.]+) s", p.stderr)]
gpu = [float(m.group(1)) for m in re.finditer(rb"GPU calls ([0-9.]+) s", p.stderr)]
print("shards %d x %d, groups %d: rc %d, wall %.2f s (%.1f fps), loop fps min %.2f max %.2f -> steady %.1f fps, lifetimes max %.2f s, gpu-call time mean %.2f s"
      % (n, k, g, p.returncode, wall, n * k / wall, min(fps), max(fps), n * k / (k / min(fps)), max(life[:-0] or [0]), sum(gpu) / max(len(gpu), 1)))
