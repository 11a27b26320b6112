"""CPU tests: pin the d>1 oracle (oracle/oracle_dn.c: q-grams, scalar aligner with the reference's
tie-breaks, greedy loop) and the d>1 host writers against golden outputs of the reference binary
(which runs its SIMD aligners), tests/golden/*.d2.* etc."""
import numpy as np
import pytest

import helpers
from helpers import GOLDEN, Oracle
from swarm_b200 import DnResult, HostDb, scoring

CASES = [("handmade", 2, False, None, "d2"), ("c1_1k_150", 2, False, None, "d2"), ("short_600_20", 2, False, None, "d2"),
         ("short_600_20", 3, False, None, "d3"), ("tie_1500_60", 2, False, None, "d2"), ("tie_1500_60", 2, True, None, "d2n"),
         ("w65_300", 3, False, None, "d3"), ("w64_400", 4, False, None, "d4"), ("w32_400", 2, False, (3, 2, 5, 3), "d2pen"),
         ("l400_250", 2, False, None, "d2"),
         # d >= 7: the reference switches to its 16-bit SIMD aligner (src/algo.cc:96-120)
         ("c1_1k_150", 7, False, None, "d7"), ("w65_300", 9, False, None, "d9"), ("w64_400", 12, False, None, "d12"),
         ("l400_250", 7, False, None, "d7"), ("handmade", 8, False, None, "d8")]


def test_scoring_conversion(built):
    # src/swarm.cc:466-483; defaults m5 p4 g12 e4 -> 18 24 13
    assert scoring() == [18, 24, 13]
    assert scoring(3, 2, 5, 3) == Oracle.scoring(None, 3, 2, 5, 3)


def test_two_gaps_vs_three_mismatches(built):
    # edit distance 2 (one deletion + one nearby insertion) but 3 differences
    # under the affine costs -> NOT linked at d=2
    a = "ACGTACGTACGTTTGACCAGTAGCATCGATCGGATTACAGGCATCGA"
    b = a[:20] + a[21:24] + "C" + a[24:]
    db = HostDb(text=f">a_2\n{a}\n>b_1\n{b}\n".encode())
    orc = Oracle(db)
    d = orc.nw_diffs(0, 1, scoring())
    ref = None
    if helpers.with_ref():
        import tempfile, os
        with tempfile.NamedTemporaryFile(suffix=".fa", delete=False) as f:
            f.write(f">a_2\n{a}\n>b_1\n{b}\n".encode())
        ref = helpers.run_ref(f.name, "-d", "2", outputs=("o",))["o"]
        os.unlink(f.name)
        assert (len(ref.splitlines()) == 1) == (d <= 2)


@pytest.mark.parametrize("name,d,ncb,pen,tag", CASES)
def test_dn_matches_reference(built, name, d, ncb, pen, tag):
    db = HostDb(GOLDEN / f"{name}.fasta", check_dup_sequences=True)
    orc = Oracle(db)
    p = scoring(*pen) if pen else scoring()
    sw, gen, par, pdiff = orc.dn_cluster(d, no_cluster_breaking=ncb, pen=p)
    res = DnResult(db, sw, gen, par, pdiff)
    assert res.swarms_text() == (GOLDEN / f"{name}.{tag}.o").read_bytes()
    assert res.stats_text() == (GOLDEN / f"{name}.{tag}.s").read_bytes()
    assert res.structure_text() == (GOLDEN / f"{name}.{tag}.i").read_bytes()
    u = GOLDEN / f"{name}.{tag}.u"
    if u.exists():                                          # -u at d>1: hits in discovery order (src/algo.cc:608-661)
        assert res.uclust_text(differences=d, penalties=p, threads=3) == u.read_bytes()
    # the oracle's own list order (rotations, src/algo.cc:205-256) is the (generation, id) order
    want = [l.split() for l in (GOLDEN / f"{name}.{tag}.o").read_text().splitlines()]
    flat = [h for l in want for h in l]
    assert [db.header(a) for a in orc.order] == flat


@pytest.mark.parametrize("name,d", [("tie_1500_60", 2), ("c1_1k_150", 2), ("w65_300", 3)])
def test_dn_writers_identical_for_any_worker_count(built, name, d):
    """the d>1 writers split their swarms over several workers like the d=1 ones: identical bytes for 1, 2, 5 workers"""
    from swarm_b200.ffi import host_lib
    db = HostDb(GOLDEN / f"{name}.fasta", check_dup_sequences=True)
    orc = Oracle(db)
    sw, gen, par, pdiff = orc.dn_cluster(d, pen=scoring())
    L = host_lib()

    def texts():
        res = DnResult(db, sw, gen, par, pdiff)
        out = [res.swarms_text(), res.stats_text(), res.structure_text()]
        res.close()
        return out

    want = texts()
    try:
        L.swbh_set_writer_grain(1)
        for workers in (1, 2, 5):
            L.swbh_set_threads(workers)
            assert texts() == want, workers
    finally:
        L.swbh_set_writer_grain(200000)
        L.swbh_set_threads(0)


@pytest.mark.parametrize("d", [20, 60, 255])
def test_wide_d_related_sequences_match_reference(built, tmp_path, d):
    """far beyond d = 6 the reference aligns with its 16-bit SIMD kernels; on RELATED sequences (variants of one seed,
    3..60 edits away) the oracle's scalar aligner reports the same difference count for every link"""
    if not helpers.have_ref():
        pytest.skip("reference binary not built")
    fa = helpers.make_variant_fasta(tmp_path / "v.fa", 500, 150, 5)
    db = HostDb(fa, check_dup_sequences=True)
    orc = Oracle(db)
    sw, gen, par, pdiff = orc.dn_cluster(d)
    res = DnResult(db, sw, gen, par, pdiff)
    r = helpers.run_ref(fa, "-d", str(d), outputs=("o", "s", "i"), threads=2)
    assert res.swarms_text() == r["o"] and res.stats_text() == r["s"] and res.structure_text() == r["i"]
