"""oracle/_ref/swarm_timed (the reference's sources with utils/progress.cc replaced by our timing shim, oracle/Makefile) must
behave exactly like oracle/_ref/swarm, the unmodified reference: the bench's CPU arm times swarm_timed.  Same output files, same
log text, on d = 1, --fastidious, d = 2 and d = 0; and the phase-time file it writes names the reference's own phases."""
import os
import subprocess

import pytest

import helpers
from helpers import GOLDEN, ROOT

TIMED = ROOT / "oracle" / "_ref" / "swarm_timed"
CASES = [("c1_1k_150", []), ("c1_1k_150", ["-f"]), ("tie_1500_60", ["-n"]), ("short_600_20", ["-d", "2"]), ("derep_mix", ["-d", "0"])]


@pytest.mark.parametrize("name,flags", CASES, ids=[f"{n} {' '.join(f)}" for n, f in CASES])
def test_timed_binary_equals_reference(built, tmp_path, name, flags):
    if not (helpers.have_ref() and TIMED.exists()):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    outs = {}
    for tag, binp in (("ref", helpers.REF_BIN), ("timed", TIMED)):
        d = tmp_path / tag
        d.mkdir()
        cmd = [str(binp), "-t", "2", "-l", str(d / "log"), "-o", str(d / "o"), "-s", str(d / "s"), "-i", str(d / "i"), "-w", str(d / "w"), *flags,
               str(GOLDEN / f"{name}.fasta")]
        env = dict(os.environ, SWARM_PHASE_TIMES=str(d / "phases"))
        p = subprocess.run(cmd, capture_output=True, env=env)
        assert p.returncode == 0, p.stderr
        outs[tag] = {k: (d / k).read_bytes() for k in ("o", "s", "i", "w", "log")}
        if tag == "timed":
            phases = [l.split("\t")[0].strip().rstrip(":") for l in (d / "phases").read_text().splitlines()]
            assert phases, "no phase times written"
            if flags == []:
                assert {"Hashing sequences", "Building network", "Clustering"} <= set(phases)
    for k in ("o", "s", "i", "w"):
        assert outs["ref"][k] == outs["timed"][k], k
    # the log differs only in its first lines (file names of this run); the text after the parameter echo is the same
    cut = lambda b: b[b.index(b"\n\n"):].replace(b"timed", b"ref")
    assert cut(outs["ref"]["log"]) == cut(outs["timed"]["log"])
