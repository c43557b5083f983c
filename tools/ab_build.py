"""Build-time variants of the library for A/B runs on the GPU box (tools/profile_run.py --lib ...).
    python tools/ab_build.py name1=-DFLAG=0 name2="-DA=1 -DB=2" ...   ->  tools/_dbg/libfastlem_b200_<name>.so
The variants are built in parallel; tools/_dbg/ is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import build as B  # noqa: E402

DBG = os.path.join(ROOT, "tools", "_dbg")


def main():
    os.makedirs(DBG, exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        out = os.path.join(DBG, f"libfastlem_b200_{name}.so")
        cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + flags.split() + ["-o", out] + B.SOURCES
        procs.append((name, out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, out, p in procs:
        log = p.communicate()[0]
        assert p.returncode == 0, log[-3000:]
        print(name, "->", out)


if __name__ == "__main__":
    main()
