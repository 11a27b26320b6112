"""CPU tests (no GPU): pin the oracle restatement (oracle/oracle_d1.c) and the host layer
(FASTA database + writers) against the golden outputs of the unmodified reference binary
committed under tests/golden/ (provenance: tests/golden/make_golden.py)."""
import numpy as np
import pytest

import helpers
from helpers import GOLDEN, Oracle
from swarm_b200 import D1Result, HostDb
from swarm_b200.ffi import network_text

CASES = ["handmade", "c1_1k_150", "tie_1500_60", "short_600_20", "w32_400", "w64_400", "w65_300"]


def test_mt19937_64_known_answer(built):
    # C++11 [rand.predef]: the 10000th consecutive invocation of a default-constructed
    # std::mt19937_64 produces 9981545732273789042.
    L = helpers.oracle_lib()
    L.orc_zobrist_exit()
    import ctypes as C
    L.orc_mt19937_64_seed.argtypes = [C.c_uint64]
    L.orc_mt19937_64_seed(5489)          # the default seed of std::mt19937_64
    vals = [L.orc_mt19937_64_next() for _ in range(10000)]
    assert vals[-1] == 9981545732273789042


def test_hashtable_size_known_answers(built):
    # /root/reference src/utils/hashtable_size.cc:61-68 (table in the source comments) + SURVEY §8
    L = helpers.oracle_lib()
    table = {11: 32, 179: 512, 2867: 8192, 45875: 131072, 734003: 2097152, 11744051: 33554432,
             187904819: 536870912, 1000: 2048, 1000000: 2097152, 10000000: 16777216, 100000000: 268435456}
    for n, want in table.items():
        assert L.orc_hashtable_size(n) == want, n


@pytest.fixture(scope="module", params=CASES)
def case(request, built):
    name = request.param
    db = HostDb(GOLDEN / f"{name}.fasta")
    orc = Oracle(db)
    assert orc.network() is not None
    orc.cluster()
    return name, db, orc


def test_variant_count_bound(case):
    name, db, orc = case
    # |V| = 6L + runs + 4 <= 7L + 4 (src/variants.cc:184-249)
    assert int(orc.net_stats[0]) <= int((7 * db.len.astype(np.int64) + 4).sum())


def test_swarms_stats_structure_match_reference(case):
    name, db, orc = case
    res = D1Result(db, orc.swarm_of, orc.generation, orc.parent)
    assert res.swarms_text() == (GOLDEN / f"{name}.o").read_bytes()
    assert res.stats_text() == (GOLDEN / f"{name}.s").read_bytes()
    assert res.structure_text() == (GOLDEN / f"{name}.i").read_bytes()
    assert res.seeds_text() == (GOLDEN / f"{name}.w").read_bytes()
    # the oracle's own linked list gives the same member order
    heads = [[db.header(a) for a in sw] for sw in orc.swarm_lists()]
    want = [l.split() for l in (GOLDEN / f"{name}.o").read_text().splitlines()]
    assert heads == want


def test_uclust_records_match_reference(case):
    """-u: C/S/H records with the scalar aligner's CIGAR (src/algod1.cc:851-934, src/nw.cc); the text must not depend
    on the number of alignment workers"""
    name, db, orc = case
    f = GOLDEN / f"{name}.u"
    if not f.exists():
        pytest.skip("no -u fixture for this case")
    res = D1Result(db, orc.swarm_of, orc.generation, orc.parent)
    assert res.uclust_text(threads=1) == f.read_bytes()
    assert res.uclust_text(threads=5) == f.read_bytes()


def test_uclust_with_grafts_and_usearch_headers(built):
    db = HostDb(GOLDEN / "c1_1k_150.fasta")
    orc = Oracle(db)
    orc.network(); orc.cluster(); orc.fastidious(boundary=3)
    res = D1Result(db, orc.swarm_of, orc.generation, orc.parent, graft_cand=orc.graft_cand, boundary=3)
    assert res.uclust_text(threads=3) == (GOLDEN / "c1_1k_150.f.u").read_bytes()
    db = HostDb(GOLDEN / "usearch_300.fasta", usearch_abundance=True)
    orc = Oracle(db)
    orc.network(); orc.cluster()
    assert D1Result(db, orc.swarm_of, orc.generation, orc.parent).uclust_text() == (GOLDEN / "usearch_300.u").read_bytes()


def test_uclust_cigar_known_answers(built):
    """run-length rules of src/utils/cigar.cc:30-60 (a count of 1 is not printed) and the trace-back priorities of
    src/nw.cc:111-191; the expected strings were produced by the reference binary (re-checked here when it is built).
    Member m against seed s: I = nucleotide only in the member, D = only in the seed."""
    seed = "ACGTTGCAAGGCTTACCGATAGGCTAACGT"
    cases = [(seed[:10] + seed[11:], "96.7", "9MD20M"),          # one G of the GG run missing: the gap goes to the first
             (seed[:10] + "T" + seed[10:], "96.8", "10MI20M"),
             (seed[:-1] + "A", "96.7", "30M"),                   # a substitution is an M column
             ("G" + seed, "96.8", "I30M"), (seed + "G", "96.8", "30MI"), (seed[2:], "93.3", "2D28M"),
             (seed[:12] + "AA" + seed[12:], "93.8", "12M2I18M"),
             (seed[:5] + "A" + seed[6:20] + seed[21:], "93.3", "20MD9M")]
    for member, pct, cigar in cases:
        text = f">s_9\n{seed}\n>m_1\n{member}\n".encode()
        db = HostDb(text=text)
        res = D1Result(db, np.zeros(2, np.uint32), np.array([0, 1], np.uint32), np.array([0xFFFFFFFF, 0], np.uint32))
        lines = res.uclust_text().decode().splitlines()
        assert lines[0] == "C\t0\t2\t*\t*\t*\t*\t*\ts_9\t*" and lines[1] == "S\t0\t30\t*\t*\t*\t*\t*\ts_9\t*"
        assert lines[2] == f"H\t0\t{len(member)}\t{pct}\t+\t0\t0\t{cigar}\tm_1\ts_9"
        if helpers.with_ref():
            import os, tempfile
            with tempfile.NamedTemporaryFile(suffix=".fa", delete=False) as f:
                f.write(text)
            ref = helpers.run_ref(f.name, "-d", "3", outputs=("u",))["u"].decode().splitlines()
            os.unlink(f.name)
            assert ref == lines


def test_uclust_band_does_not_change_the_records(built, tmp_path):
    """the -u aligner fills a band that doubles until the result is provably the full matrix's: every first band
    (1, 2, 8, 64) must give the records of the full matrix (0) — members from 1 to 60 edits away from their seed,
    unrelated sequences of other lengths, very short ones"""
    import random
    from swarm_b200.ffi import host_lib
    rng = random.Random(11)
    fa = helpers.make_variant_fasta(tmp_path / "v.fa", 300, 120, 8, kmin=1, kmax=60)
    text = fa.read_text()
    for k in range(60):                                    # unrelated sequences, lengths 1..260
        L = rng.choice([1, 2, 3, 17, 64, 119, 120, 121, 260])
        text += f">x{k}_1\n{''.join(rng.choice('ACGT') for _ in range(L))}\n"
    db = HostDb(text=text.encode())
    n = db.n
    # one swarm: the most abundant amplicon is the seed of everything (alignment partners are then seed vs. member)
    sw = np.zeros(n, np.uint32)
    gen = np.ones(n, np.uint32); gen[0] = 0
    par = np.zeros(n, np.uint32); par[0] = 0xFFFFFFFF
    res = D1Result(db, sw, gen, par)
    L = host_lib()
    try:
        L.swbh_uclust_band(0)
        want = res.uclust_text(threads=4)
        for w in (1, 2, 8, 64):
            L.swbh_uclust_band(w)
            assert res.uclust_text(threads=4) == want, w
    finally:
        L.swbh_uclust_band(8)
    assert want.count(b"\nH\t") == n - 1


def _all_texts(db, orc, pairs):
    rp = np.zeros(db.n + 1, dtype=np.uint64)
    np.add.at(rp, pairs[:, 0].astype(np.int64) + 1, 1)
    rp = np.cumsum(rp).astype(np.uint64)
    out = [network_text(db, rp, pairs[:, 1])]
    for gc in (None, orc.graft_cand):
        res = D1Result(db, orc.swarm_of, orc.generation, orc.parent, graft_cand=gc, boundary=3)
        out += [res.swarms_text(), res.swarms_text(mothur=True), res.stats_text(), res.structure_text(), res.seeds_text()]
        res.close()
    return out


def test_writers_identical_for_any_worker_count(case):
    """the output writers cut their swarms / rows into ranges of equal work for several workers (host/d1_result.cc: parallel_text):
    identical bytes for 1, 2, 3 and 7 workers, with the grain lowered so that even these small inputs are split — including ranges
    that start inside runs of attached (grafted) swarms, whose output numbers come from a prefix count"""
    from swarm_b200.ffi import host_lib
    name, db, _ = case
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    orc.fastidious(boundary=3)
    pairs = orc.links()
    L = host_lib()
    want = _all_texts(db, orc, pairs)
    assert want[1] == (GOLDEN / f"{name}.o").read_bytes() and want[6] == (GOLDEN / f"{name}.f.o").read_bytes()
    try:
        L.swbh_set_writer_grain(1)
        for workers in (1, 2, 3, 7):
            L.swbh_set_threads(workers)
            assert _all_texts(db, orc, pairs) == want, workers
    finally:
        L.swbh_set_writer_grain(200000)
        L.swbh_set_threads(0)


def test_writers_with_default_grain_on_300k(built, tmp_path):
    """above the default grain the writers really run on several workers: same bytes as one worker"""
    import helpers
    from swarm_b200.ffi import host_lib
    db = HostDb(helpers.make_fasta(tmp_path / "w.fa", 300000, 120, 8, 1))
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    orc.fastidious(boundary=3)
    pairs = orc.links()
    L = host_lib()
    try:
        L.swbh_set_threads(1)
        want = _all_texts(db, orc, pairs)
        L.swbh_set_threads(6)
        assert _all_texts(db, orc, pairs) == want
    finally:
        L.swbh_set_threads(0)


def test_oracle_and_parallel_writers_reproduce_the_reference_digests_at_1M(built, tmp_path):
    """1 M x 150 bp (the generator stream of the bench, seed 42): the CPU oracle's clustering, written by the host writers on 8
    workers, hashes to the digests of the reference's own -o / -s / -i files (tests/golden/scale_hashes.json[c2_1m]) — pins the
    oracle and the multi-worker writers at a size where every range boundary falls inside real data"""
    from swarm_b200.ffi import host_lib
    db = HostDb(helpers.make_fasta(tmp_path / "m.fa", 1000000, 150, 42, 0))
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    L = host_lib()
    try:
        L.swbh_set_threads(8)
        res = D1Result(db, orc.swarm_of, orc.generation, orc.parent)
        got = helpers.output_hashes(res.swarms_text(), res.stats_text(), res.structure_text())
        res.close()
    finally:
        L.swbh_set_threads(0)
    assert helpers.compare_case("c2_1m", got) == []


def test_network_matches_reference(case):
    name, db, orc = case
    pairs = orc.links()
    rp = np.zeros(db.n + 1, dtype=np.uint64)
    np.add.at(rp, pairs[:, 0].astype(np.int64) + 1, 1)
    rp = np.cumsum(rp).astype(np.uint64)
    assert network_text(db, rp, pairs[:, 1]) == (GOLDEN / f"{name}.j").read_bytes()


def test_no_cluster_breaking(case):
    name, db, _ = case
    orc = Oracle(db)
    orc.network(no_cluster_breaking=True)
    orc.cluster()
    res = D1Result(db, orc.swarm_of, orc.generation, orc.parent)
    assert res.swarms_text() == (GOLDEN / f"{name}.n.o").read_bytes()


@pytest.mark.parametrize("boundary,suffix", [(3, "f"), (10, "f.b10")])
def test_fastidious_matches_reference(case, boundary, suffix):
    name, db, _ = case
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    orc.fastidious(boundary=boundary)
    res = D1Result(db, orc.swarm_of, orc.generation, orc.parent, graft_cand=orc.graft_cand, boundary=boundary)
    assert res.swarms_text() == (GOLDEN / f"{name}.{suffix}.o").read_bytes()
    if suffix == "f":
        assert res.stats_text() == (GOLDEN / f"{name}.f.s").read_bytes()
        assert res.structure_text() == (GOLDEN / f"{name}.f.i").read_bytes()
    heads = [[db.header(a) for a in sw] for sw in orc.swarm_lists()]
    want = [l.split() for l in (GOLDEN / f"{name}.{suffix}.o").read_text().splitlines()]
    assert heads == want


def test_usearch_and_mothur(built):
    db = HostDb(GOLDEN / "usearch_300.fasta", usearch_abundance=True)
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    res = D1Result(db, orc.swarm_of, orc.generation, orc.parent)
    assert res.swarms_text() == (GOLDEN / "usearch_300.o").read_bytes()
    assert res.stats_text() == (GOLDEN / "usearch_300.s").read_bytes()
    assert res.structure_text() == (GOLDEN / "usearch_300.i").read_bytes()
    assert res.seeds_text() == (GOLDEN / "usearch_300.w").read_bytes()
    assert res.swarms_text(mothur=True) == (GOLDEN / "usearch_300.r.o").read_bytes()


def test_duplicate_sequences_detected(built):
    db = HostDb(text=b">a_3\nACGTACGT\n>b_2\nACGTACGT\n>c_1\nACGTACGA\n")
    assert Oracle(db).network() is None     # the reference aborts (src/algod1.cc:1141-1150)


@pytest.mark.parametrize("text,msg", [
    (b"ACGT\n", "Illegal header line in fasta file."),
    (b">a_1\nACGN\n", "Illegal character 'N' in sequence on line 2."),
    (b">a_1\n>b_1\nAC\n", "Empty sequence found on line 1."),
    (b">a\nACGT\n", "Abundance annotations not found for 1 sequences, starting on line 1."),
    (b">a_0\nACGT\n", "Illegal abundance value on line 1:"),
    (b">a_1\nACGT\n>a_2\nACGA\n", "Duplicated sequence identifier: a"),
    (b">_1\nACGT\n", "Empty sequence identifier."),
])
def test_fasta_errors(built, text, msg):
    with pytest.raises(ValueError) as e:
        HostDb(text=text)
    assert msg in str(e.value)
    if helpers.with_ref():      # the reference prints the same message (src/db.cc)
        import tempfile, os
        with tempfile.NamedTemporaryFile(suffix=".fa", delete=False) as f:
            f.write(text)
        r = helpers.run_ref(f.name)
        os.unlink(f.name)
        assert r["rc"] == 1 and msg.encode() in r["stderr"]


def _ed_le2(a, b):
    """banded Levenshtein: True iff 1 <= ed(a, b) <= 2"""
    la, lb = len(a), len(b)
    if abs(la - lb) > 2:
        return False
    prev = {c: c + 1 for c in range(-1, min(lb, 2))}
    for r in range(la):
        cur = {}
        if r - 2 <= -1:
            cur[-1] = r + 1
        for c in range(max(0, r - 2), min(lb, r + 3)):
            best = 99
            if c - 1 in prev:
                best = prev[c - 1] + (a[r] != b[c])
            if c in prev:
                best = min(best, prev[c] + 1)
            if c - 1 in cur:
                best = min(best, cur[c - 1] + 1)
            cur[c] = best
        if min(cur.values()) > 2:
            return False
        prev = cur
    d = prev.get(lb - 1, 99)
    return 1 <= d <= 2


@pytest.mark.parametrize("name", ["handmade", "short_600_20", "w32_400"])
def test_fastidious_is_edit_distance_two(built, name):
    """the closed form the GPU join uses: graft_cand[l] = min heavy h with 1 <= ed(h,l) <= 2 equals the
    reference's two-level microvariant scheme (oracle restatement of src/algod1.cc:374-552)"""
    db = HostDb(GOLDEN / f"{name}.fasta")
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    if orc.fastidious(boundary=3) < 0:
        pytest.skip("only light or only heavy swarms")
    seqs = []
    for i in range(db.n):
        w = db.words[i * db.stride:(i + 1) * db.stride]
        p = np.arange(int(db.len[i]))
        seqs.append(((w[p >> 5] >> ((p & 31).astype(np.uint64) << np.uint64(1))) & np.uint64(3)).astype(np.int8).tolist())
    mass = orc.sw_mass_before if hasattr(orc, "sw_mass_before") else None
    o2 = Oracle(db); o2.network(); o2.cluster()
    light = o2.sw_mass[o2.swarmid] < 3
    want = np.full(db.n, 0xFFFFFFFF, dtype=np.uint32)
    heavy_ids = np.nonzero(~light)[0]
    for l in np.nonzero(light)[0]:
        for h in heavy_ids:
            if _ed_le2(seqs[h], seqs[l]):
                want[l] = h
                break
    assert np.array_equal(want, orc.graft_raw)
