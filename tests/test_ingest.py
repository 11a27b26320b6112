"""FASTA ingest: the multi-threaded parser of swarm_b200/host/amplicon_db.cc must build exactly
the database of the serial one (which mirrors /root/reference src/db.cc and is pinned by the golden tests), and every
malformed input must end with the serial parser's — i.e. the reference's — message."""
import os
import random

import numpy as np
import pytest

import helpers
from helpers import GOLDEN
from swarm_b200 import HostDb
from swarm_b200.ffi import host_lib


def _parse(text, threads, **kw):
    L = host_lib()
    os.environ["SWARM_B200_INGEST_MIN_BYTES"] = "1"          # test hook: small inputs take the parallel path too
    L.swbh_set_threads(threads)
    try:
        try:
            db = HostDb(text=text, **kw)
        except ValueError as e:
            return ("error", str(e))
        out = ("ok", db.n, db.longest, db.stride, db.nucleotides, db.words.tobytes(), db.len.tobytes(), db.abundance.tobytes(),
               tuple(db.headers()), db.opts)
        db.close()
        return out
    finally:
        L.swbh_set_threads(0)
        os.environ.pop("SWARM_B200_INGEST_MIN_BYTES", None)


FASTAS = sorted(p.name for p in GOLDEN.glob("*.fasta"))


@pytest.mark.parametrize("name", FASTAS)
def test_parallel_equals_serial_on_golden_inputs(built, name):
    text = (GOLDEN / name).read_bytes()
    kw = {"usearch_abundance": name.startswith("usearch") or name.endswith("_z.fasta")}
    want = _parse(text, 1, **kw)
    assert want[0] == "ok"
    for t in (2, 3, 8):
        assert _parse(text, t, **kw) == want
    # d > 1 adds the duplicate-sequence check (src/db.cc:763-796): same verdict from both parsers
    assert _parse(text, 5, check_dup_sequences=True, **kw) == _parse(text, 1, check_dup_sequences=True, **kw)


BAD = [
    b">a_1\nACGT\n>b_2\nACXT\n",                                  # illegal character
    b">a_1\nACGT\n>b_2\nAC\x01T\n",                               # illegal non-printable character
    b">a_1\nACGT\n>b\nACGT\n>c\nAAAA\n",                          # abundance annotations missing
    b">a_1\nACGT\n>a_2\nACGA\n",                                  # duplicated identifier
    b">a_1\nACGT\n>b_0\nACGA\n",                                  # illegal abundance
    b">a_1\nACGT\n>_3\nACGA\n",                                   # empty identifier
    b">a_1\n>b_2\nACGT\n",                                        # empty sequence
    b"ACGT\n>b_2\nACGT\n",                                        # no header first
    b"\n>a_1\nACGT\n",
    b">a_1\nAC\x00GT\n>b_1\nAAAA\n",                              # NUL byte (C-string semantics of the reference's reader)
    b">a_1\nACGT\n>b_1\nAC>GT\n",                                 # '>' inside a sequence line
]


@pytest.mark.parametrize("i", range(len(BAD)))
def test_malformed_inputs_give_the_serial_message(built, i):
    pad = b"".join(b">p%d_3\nACGTACGTAC\n" % k for k in range(40))       # several ranges, the defect in the middle
    for text in (BAD[i], pad + BAD[i] + pad.replace(b">p", b">q")):
        want = _parse(text, 1)
        for t in (2, 7):
            assert _parse(text, t) == want
    assert _parse(BAD[i], 1)[0] in ("error", "ok")


def test_duplicate_sequences_only_matter_when_asked(built):
    text = b">a_3\nACGT\n>b_2\nacgu\n>c_1\nAC\nGT\n"
    assert _parse(text, 4)[0] == "ok" and _parse(text, 4) == _parse(text, 1)
    e1, e4 = _parse(text, 1, check_dup_sequences=True), _parse(text, 4, check_dup_sequences=True)
    assert e1 == e4 and e1[0] == "error" and "identical sequences" in e1[1]
    # equal packed words, different lengths: not duplicates
    ok = b">a_3\nA\n>b_2\nAA\n>c_1\nAAA\n"
    assert _parse(ok, 3, check_dup_sequences=True) == _parse(ok, 1, check_dup_sequences=True) and _parse(ok, 3, check_dup_sequences=True)[0] == "ok"


@pytest.mark.parametrize("seed", range(6))
def test_random_layouts_and_range_boundaries(built, seed):
    """records with wrapped sequence lines, CRLF, blank lines, '>' and spaces inside header lines, no final newline:
    every split of the text into worker ranges must give the serial database"""
    rng = random.Random(seed)
    recs = []
    for k in range(rng.randint(30, 400)):
        L = rng.choice([1, 2, 31, 32, 33, 64, 65, rng.randint(1, 200)])
        s = "".join(rng.choice("ACGTacgtUu") for _ in range(L))
        w = rng.choice([L, 7, 60, 1])
        nl = rng.choice(["\n", "\n", "\r\n"])
        lines = nl.join(s[i:i + w] for i in range(0, L, w))
        if rng.random() < 0.1:
            lines += nl                                              # a blank line inside the record
        hdr = f"x{k}{rng.choice(['', 'y>z', '-'])}_{rng.choice([1, 1, 2, 9, 9, 500])}{rng.choice(['', ' desc > more', ' t'])}"
        recs.append(f">{hdr}{nl}{lines}{nl}")
    text = "".join(recs)
    if seed % 2:
        text = text.rstrip("\r\n")
    text = text.encode()
    want = _parse(text, 1)
    assert want[0] == "ok"
    for t in (2, 3, 5, 8, 13):
        assert _parse(text, t) == want


def test_large_input_takes_the_parallel_path_by_default(built, tmp_path):
    fa = tmp_path / "big.fa"
    helpers.make_fasta(fa, 20000, 100, 3)                             # > 1 MiB: parallel without the test hook
    text = fa.read_bytes()
    assert len(text) > (1 << 20)
    L = host_lib()
    L.swbh_set_threads(1)
    a = HostDb(fa)
    L.swbh_set_threads(0)
    b = HostDb(fa)
    assert a.n == b.n == 20000 and np.array_equal(a.words, b.words) and np.array_equal(a.len, b.len)
    assert np.array_equal(a.abundance, b.abundance) and a.headers() == b.headers()


def test_compact_form_of_the_database(built):
    """swbh_db_compact (host C ABI) == ffi.compact_form (numpy): u16 lengths + abundance runs that expand to the arrays"""
    import ctypes as C
    from swarm_b200.ffi import compact_form
    L = host_lib()
    L.swbh_db_compact.restype = C.c_uint32
    L.swbh_db_compact.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint32))]
    for name in ("tie_1500_60", "c1_1k_150", "handmade"):
        db = HostDb(GOLDEN / f"{name}.fasta")
        p16, pab, pst = C.POINTER(C.c_uint16)(), C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)()
        runs = L.swbh_db_compact(db._h, C.byref(p16), C.byref(pab), C.byref(pst))
        l16, rab, rst = compact_form(db.len, db.abundance)
        assert runs == len(rab)
        assert np.array_equal(np.ctypeslib.as_array(p16, shape=(db.n,)), l16)
        assert np.array_equal(np.ctypeslib.as_array(pab, shape=(runs,)), rab)
        assert np.array_equal(np.ctypeslib.as_array(pst, shape=(runs + 1,)), rst)
        assert np.array_equal(np.repeat(rab, np.diff(rst.astype(np.int64))), db.abundance)


def test_extreme_shapes(built):
    """ranges that hold no record start (one 3 MB sequence wrapped at 80 columns), more workers than records, 200 k
    one-nucleotide records, append-abundance (-a) with and without annotations: parallel == serial"""
    rng = random.Random(4)
    big = "".join(rng.choice("ACGT") for _ in range(3_000_000))
    wrapped = "\n".join(big[i:i + 80] for i in range(0, len(big), 80))
    cases = [
        (f">big_7\n{wrapped}\n>small_3\nACGT\n".encode(), {}),
        (f">only_1\n{big[:500]}".encode(), {}),
        ("".join(f">t{i}_{1 + i % 3}\n{'ACGT'[i % 4]}{'ACGT'[(i // 4) % 4]}{'ACGT'[(i // 16) % 4]}\n" for i in range(200_000)).encode(), {}),
        (b">a\nACGT\n>b_5\nACGA\n>c;size=3;\nAAAA\n", {"append_abundance": 2}),
        (b">a;size=4;\nACGT\n>b\nACGA\n", {"append_abundance": 9, "usearch_abundance": True}),
    ]
    for text, kw in cases:
        want = _parse(text, 1, **kw)
        assert want[0] == "ok", want
        for t in (2, 8, 32):
            assert _parse(text, t, **kw) == want
        assert _parse(text, 6, check_dup_sequences=True, **kw) == _parse(text, 1, check_dup_sequences=True, **kw)
