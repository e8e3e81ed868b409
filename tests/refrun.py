"""Helpers that drive the compiled reference (oracle/_ref, test infrastructure) and the golden fixtures."""
import lzma
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_ref(name="x264_dump"):
    return os.path.exists(os.path.join(REF_DIR, name))


def synth_clip(pcamv, width, height, frames, config=1, stream=0, noise16=32, workdir=None):
    exe = pcamv.build.build_synth()
    workdir = workdir or tempfile.mkdtemp(prefix="pcamv_")
    path = os.path.join(workdir, "clip_%dx%d_%d_%d_%d_%d.yuv" % (width, height, frames, config, stream, noise16))
    if not os.path.exists(path):
        subprocess.check_call([exe, str(width), str(height), str(frames), str(config), str(stream), path, str(noise16)])
    return path


def run_ref(clip, width, height, args, binary="x264_dump", dump=None, frames=None, planes=True, calls=True,
            stats=None, count=False, out=None, extra_env=None, timeout=3600):
    """Run the reference CLI; returns (bitstream path, dump path or None)."""
    workdir = os.path.dirname(clip)
    out = out or os.path.join(workdir, "out_%s.264" % binary)
    env = dict(os.environ)
    if dump:
        env["PCAMV_DUMP"] = dump
        env["PCAMV_DUMP_PLANES"] = "1" if planes else "0"
        env["PCAMV_DUMP_CALLS"] = "1" if calls else "0"
        if frames:
            env["PCAMV_DUMP_FRAMES"] = frames
    if stats:
        env["PCAMV_STATS"] = stats
    if count:
        env["PCAMV_COUNT"] = "1"
    env.update(extra_env or {})
    cmd = [os.path.join(REF_DIR, binary)] + list(args) + ["-o", out, clip, "%dx%d" % (width, height)]
    subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True, timeout=timeout)
    return out, dump


def golden_dump_path(name, workdir=None):
    """Decompress tests/golden/<name>.bin.xz into a temp file and return its path."""
    src = os.path.join(GOLDEN, name + ".bin.xz")
    workdir = workdir or tempfile.mkdtemp(prefix="pcamv_golden_")
    dst = os.path.join(workdir, name + ".bin")
    if not os.path.exists(dst):
        with lzma.open(src, "rb") as f, open(dst, "wb") as g:
            g.write(f.read())
    return dst
