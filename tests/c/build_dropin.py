"""Builds tests/c/_build/dropin: dropin.c compiled against the reference's include/msb_64.h
(verbatim, from /root/reference) and linked against libmsb64_b200.so.  Runs where
/root/reference exists (this container); the GPU box uses the prebuilt binary."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_INCLUDE = "/root/reference/include"
OUT = os.path.join(HERE, "_build", "dropin")
LIB_DIR = os.path.join(ROOT, "inplacemsdradixsort_b200", "lib")


def build() -> str | None:
    """Returns the binary's path, or None when the reference header is not available here."""
    if not os.path.exists(os.path.join(REF_INCLUDE, "msb_64.h")):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["gcc", "-O2", "-Wall", "-I", REF_INCLUDE, os.path.join(HERE, "dropin.c"), "-o", OUT,
           "-L", LIB_DIR, "-lmsb64_b200", "-Wl,-rpath,$ORIGIN/../../../inplacemsdradixsort_b200/lib"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
