#!/usr/bin/env python3
"""Build a variant of libsept_b200.so next to the shipped one (A/B timing with tools/ab.sh, diagnostic builds):

    SEPT_NVCC_EXTRA="-DFOO=1" python tools/build_variant.py variants/foo.so
"""
import os
import shutil
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_emotion_privacy_trust_b200 import build

dst = Path(sys.argv[1])
dst.parent.mkdir(parents=True, exist_ok=True)
keep = build.LIB.with_suffix(".so.keep")
if build.LIB.exists():
    shutil.copyfile(build.LIB, keep)
try:
    build.build(force=True)
    shutil.copyfile(build.LIB, dst)
finally:
    if keep.exists():
        os.replace(keep, build.LIB)
print(dst, "built with", os.environ.get("SEPT_NVCC_EXTRA", "(no extra flags)"))
