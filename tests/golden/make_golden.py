#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/swarm, built by oracle/Makefile from /root/reference) on small inputs.

    python tests/golden/make_golden.py          # only possible where oracle/_ref/swarm exists

Each case <name>.fasta gets: <name>.o (swarms), .s (stats), .i (structure), .j (network) at d=1;
<name>.u (UCLUST-like records) for the cases of UCLUST below; <name>.n.o with -n; <name>.f.o/.f.s/.f.i with --fastidious (-t 1: the reference's light pass has
unsynchronised inserts).  Inputs come from tools/gen_amplicons.c (seeded) or are
hand-made below.  The fixtures are committed; this script is their provenance.
"""
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import helpers  # noqa: E402

HANDMADE = b""">seedA_100
ACGTACGTACGTACGTACGTACGTACGTACGTAAAACCCCGGGGTTTT
>subB_40
ACGTACGTACGTACGTACGTACGTACGTACGTAAAACCCCGGGGTTTA
>delC_12
ACGTACGTACGTACGTACGTACGTACGTACGTAAACCCCGGGGTTTT
>insD_12
ACGTACGTACGTACGTACGTACGTACGTACGTAAAACCCCGGGGGTTTT
>lower_u_7
acguacguacguacguacguacguacguacguaaaaccccgggguuca
>gen2_3
ACGTACGTACGTACGTACGTACGTACGTACGTAAACCCCGGGGTTTA
>tieX_5
TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT
>tieY_5
TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT
>tieZ_5
TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTGTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT
>crlf_2\r
GGGGGGGGCCCCCCCC\r
AAAATTTT\r
>orph_1
ACGTACGTACGTACGTACGTACGTACGTACGTAAAACCCCGGGGTTAA
>orph2_1
ACGTACGTACGTACGTACGTACGTACGTACGTAAAACCCCGGGTTTAA
>lonely_1
CATCATCATCATCATCATCAT
>one_9
A
>two_4
AC
>three_1
C
"""

CASES = [
    # name, n, L, seed, ab_mode, orphan_p
    ("c1_1k_150", 1000, 150, 42, 0, 0.2),
    ("tie_1500_60", 1500, 60, 7, 1, 0.1),
    ("short_600_20", 600, 20, 3, 0, 0.2),
    ("w32_400", 400, 32, 5, 0, 0.2),
    ("w64_400", 400, 64, 6, 1, 0.2),
    ("w65_300", 300, 65, 8, 0, 0.3),
]


# -u (UCLUST-like records: alignments + CIGAR): name, flags, tag
UCLUST = [("handmade", [], ""), ("c1_1k_150", [], ""), ("c1_1k_150", ["-f"], "f."), ("short_600_20", [], ""),
          ("w65_300", [], ""), ("handmade", ["-d", "2"], "d2."), ("short_600_20", ["-d", "3"], "d3."),
          ("tie_1500_60", ["-d", "2", "-n"], "d2n."), ("w32_400", ["-d", "2", "-m", "3", "-p", "2", "-g", "5", "-e", "3"], "d2pen."),
          ("l400_250", ["-d", "2"], "d2."), ("usearch_300", ["-z"], "")]


def make_uclust():
    for name, flags, tag in UCLUST:
        r = helpers.run_ref(HERE / f"{name}.fasta", *flags, outputs=("u",), threads=1)
        assert r["rc"] == 0, r["stderr"]
        (HERE / f"{name}.{tag}u").write_bytes(r["u"])
    print("uclust ok")


# d >= 7 with the default scoring: the reference switches to its 16-bit aligner (src/algo.cc:96-120), the engine to
# k_dn_align_wide (band half-width > 15)
DN_WIDE = [("c1_1k_150", 7), ("w65_300", 9), ("w64_400", 12), ("l400_250", 7), ("handmade", 8)]


def make_dn_wide():
    for name, d in DN_WIDE:
        r = helpers.run_ref(HERE / f"{name}.fasta", "-d", str(d), outputs=("o", "s", "i"), threads=2)
        assert r["rc"] == 0, r["stderr"]
        for k in "osi":
            (HERE / f"{name}.d{d}.{k}").write_bytes(r[k])
    print("d>=7 ok")


def make_derep():
    """d=0 inputs need identical sequences: reads drawn (seeded) from the first 250 sequences of c1_1k_150 with fresh
    labels and small abundances (many 1s -> singletons, equal masses -> the seed-index tie-break), some in lower
    case / with U, plus sequences that pack to the same words and differ only in length (A, AA, AAA)."""
    import random
    rng = random.Random(20261017)
    seqs = [l for l in (HERE / "c1_1k_150.fasta").read_text().splitlines() if not l.startswith(">")][:250]
    reads = []
    for i, sq in enumerate(seqs):
        for k in range(rng.choice([1, 1, 1, 2, 2, 3, 4, 6, 9])):
            t = sq
            r = rng.random()
            if r < 0.15:
                t = sq.lower()
            elif r < 0.3:
                t = sq.replace("T", "U")
            reads.append((f"r{i}x{k}", rng.choice([1, 1, 1, 2, 3, 5, 40]), t))
    for k, t in enumerate(["A", "AA", "AAA", "a", "AA", "C", "ACGT" * 16, "ACGT" * 16 + "A", "ACGT" * 16]):
        reads.append((f"tiny{k}", rng.choice([1, 2, 2]), t))
    rng.shuffle(reads)
    (HERE / "derep_mix.fasta").write_text("".join(f">{h}_{ab}\n{t}\n" for h, ab, t in reads))
    (HERE / "derep_mix_z.fasta").write_text("".join(f">{h};size={ab};\n{t}\n" for h, ab, t in reads[:300]))
    for name, flags in [("derep_mix", []), ("derep_mix_z", ["-z"])]:
        r = helpers.run_ref(HERE / f"{name}.fasta", "-d", "0", *flags, outputs=("o", "s", "i", "w", "u"))
        assert r["rc"] == 0, r["stderr"]
        for k in "osiwu":
            (HERE / f"{name}.d0.{k}").write_bytes(r[k])
        (HERE / f"{name}.d0.log").write_bytes(r["log"][r["log"].index(b"\nNumber of swarms"):])
        r = helpers.run_ref(HERE / f"{name}.fasta", "-d", "0", "-r", *flags, outputs=("o",))
        (HERE / f"{name}.d0.r.o").write_bytes(r["o"])
    print("derep ok")


def main():
    if not helpers.have_ref():
        sys.exit("oracle/_ref/swarm missing: run `make -C oracle ref` where /root/reference exists")
    if sys.argv[1:] == ["uclust"]:          # only (re)generate the -u fixtures; the inputs must exist
        return make_uclust()
    if sys.argv[1:] == ["derep"]:
        return make_derep()
    if sys.argv[1:] == ["dn_wide"]:
        return make_dn_wide()
    (HERE / "handmade.fasta").write_bytes(HANDMADE)
    names = ["handmade"]
    for name, n, L, seed, mode, op in CASES:
        helpers.make_fasta(HERE / f"{name}.fasta", n, L, seed, mode, op)
        names.append(name)
    for name in names:
        fa = HERE / f"{name}.fasta"
        r = helpers.run_ref(fa, outputs=("o", "s", "i", "j", "w"))
        assert r["rc"] == 0, r["stderr"]
        for k in "osijw":
            (HERE / f"{name}.{k}").write_bytes(r[k])
        r = helpers.run_ref(fa, "-n", outputs=("o",))
        (HERE / f"{name}.n.o").write_bytes(r["o"])
        r = helpers.run_ref(fa, "-f", outputs=("o", "s", "i"), threads=1)
        assert r["rc"] == 0, r["stderr"]
        for k in "osi":
            (HERE / f"{name}.f.{k}").write_bytes(r[k])
        r = helpers.run_ref(fa, "-f", "-b", "10", outputs=("o",), threads=1)
        (HERE / f"{name}.f.b10.o").write_bytes(r["o"])
        print(name, "ok")
    # d > 1 (the reference runs its SIMD aligners here): swarms, stats, structure
    for name, flags, tag in [("handmade", ["-d", "2"], "d2"), ("c1_1k_150", ["-d", "2"], "d2"), ("short_600_20", ["-d", "2"], "d2"),
                             ("short_600_20", ["-d", "3"], "d3"), ("tie_1500_60", ["-d", "2"], "d2"), ("tie_1500_60", ["-d", "2", "-n"], "d2n"),
                             ("w65_300", ["-d", "3"], "d3"), ("w64_400", ["-d", "4"], "d4"),
                             ("w32_400", ["-d", "2", "-m", "3", "-p", "2", "-g", "5", "-e", "3"], "d2pen")]:
        r = helpers.run_ref(HERE / f"{name}.fasta", *flags, outputs=("o", "s", "i"), threads=2)
        assert r["rc"] == 0, r["stderr"]
        for k in "osi":
            (HERE / f"{name}.{tag}.{k}").write_bytes(r[k])
    make_dn_wide()
    helpers.make_fasta(HERE / "l400_250.fasta", 250, 400, 12, 0, 0.3)
    r = helpers.run_ref(HERE / "l400_250.fasta", "-d", "2", outputs=("o", "s", "i"), threads=2)
    for k in "osi":
        (HERE / f"l400_250.d2.{k}").write_bytes(r[k])
    print("d>1 ok")
    # usearch-style headers (-z) + mothur (-r): derived from c1 by rewriting headers
    src = (HERE / "c1_1k_150.fasta").read_bytes().splitlines()
    out = []
    for ln in src[:600]:
        if ln.startswith(b">"):
            lab, ab = ln[1:].rsplit(b"_", 1)
            out.append(b">" + lab + b";size=" + ab + b";")
        else:
            out.append(ln)
    (HERE / "usearch_300.fasta").write_bytes(b"\n".join(out) + b"\n")
    r = helpers.run_ref(HERE / "usearch_300.fasta", "-z", outputs=("o", "s", "i", "w"))
    assert r["rc"] == 0, r["stderr"]
    for k in "osiw":
        (HERE / f"usearch_300.{k}").write_bytes(r[k])
    r = helpers.run_ref(HERE / "usearch_300.fasta", "-z", "-r", outputs=("o",))
    (HERE / "usearch_300.r.o").write_bytes(r["o"])
    print("usearch ok")
    make_uclust()
    make_derep()


if __name__ == "__main__":
    main()


# ---- log text (from "Reading sequences" on) of the reference run with -l, for the command-line tests: see LOG_CASES in tests/test_cli.py
def make_logs():
    import os
    import subprocess
    import tempfile
    from test_cli import LOG_CASES
    for name, flags, tag in LOG_CASES:
        with tempfile.TemporaryDirectory() as td:
            fl = [os.path.join(td, f) if len(f) == 1 and f.isupper() else f for f in flags]
            subprocess.run([str(helpers.REF_BIN), "-l", os.path.join(td, "log"), *fl, str(HERE / f"{name}.fasta")], check=True, capture_output=True)
            log = open(os.path.join(td, "log"), "rb").read()
            (HERE / f"{name}.{tag}").write_bytes(log[log.index(b"Reading sequences"):])


if __name__ == "__main__" and "logs" in sys.argv[1:]:
    make_logs()
