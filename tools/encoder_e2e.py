#!/usr/bin/env python3
"""Whole-encoder legs of bench.py on their own (single stream and GOP-sharded), for quick iteration."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import pcamv_loader  # noqa: E402

pcamv = pcamv_loader.load()
wd = tempfile.mkdtemp(prefix="pcamv_enc_")
print(json.dumps(bench.encoder_e2e(pcamv, wd, 0)))
print(json.dumps(bench.encoder_e2e_sharded(pcamv, wd, 0)))
