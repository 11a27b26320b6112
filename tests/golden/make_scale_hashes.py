#!/usr/bin/env python
"""Provenance of tests/golden/scale_hashes.json: the UNMODIFIED reference (oracle/_ref/swarm, compiled from
/root/reference by oracle/Makefile) run on the seeded synthetic sets of tests/helpers.py::SCALE_CASES at their full
BASELINE sizes; only SHA-256 digests of its -o (raw and canonical), -s and -i files are kept.

    python tests/golden/make_scale_hashes.py [case ...]        # default: every case not yet in the file
"""
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402


def main():
    assert helpers.have_ref(), "oracle/_ref/swarm missing: run `make oracle` where /root/reference exists"
    done = helpers.scale_hashes()
    names = sys.argv[1:] or [k for k in helpers.SCALE_CASES if k not in done]
    for name in names:
        n, L, seed, ab_mode, flags, threads = helpers.SCALE_CASES[name]
        fa = helpers.scale_fasta(name)
        with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
            out = {k: os.path.join(td, k) for k in "osi"}
            cmd = [str(helpers.REF_BIN), "-t", str(threads), "-l", os.devnull, "-o", out["o"], "-s", out["s"], "-i", out["i"], *flags, fa]
            t0 = time.time()
            subprocess.run(cmd, check=True)
            dt = time.time() - t0
            h = helpers.output_hashes(*(open(out[k], "rb").read() for k in "osi"))
        h.update({"n": n, "length": L, "seed": seed, "ab_mode": ab_mode, "flags": " ".join(flags), "reference_threads": threads,
                  "reference_wall_s": round(dt, 1), "reference": "torognes/swarm 3.1.6, oracle/_ref/swarm"})
        done = helpers.scale_hashes()
        done[name] = h
        helpers.SCALE_HASHES.write_text(json.dumps(done, indent=1, sort_keys=True) + "\n")
        print(name, h, flush=True)


if __name__ == "__main__":
    main()
