"""Build the C restatement of the oracle (oracle/kalman_c.c) into oracle/_build/libkalman_c.so with gcc + OpenMP.

TEST INFRASTRUCTURE ONLY.  `__graft_entry__.build()` calls this so that the built checker travels to the GPU box."""

from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kalman_c.c")
OUT_DIR = os.path.join(HERE, "_build")


def _cpu_tag() -> str:
    """-march=native code only runs on the CPU model it was built for: key the artefact by the model name so the
    GPU box rebuilds instead of loading a foreign binary."""
    import hashlib

    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("model name", "flags")):
                    model += line
                    if line.startswith("flags"):
                        break
    except OSError:
        pass
    return hashlib.md5(model.encode()).hexdigest()[:10]


LIB = os.path.join(OUT_DIR, f"libkalman_c_{_cpu_tag()}.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:  # e.g. -march=native unsupported
        cmd.remove("-march=native")
        res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
