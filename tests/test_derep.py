"""d = 0 (dereplication, /root/reference src/derep.cc): the oracle restatement (oracle/oracle_d0.c) and the host
writers against golden outputs of the reference binary (CPU), and the engine's kernels (csrc/d0_derep.cuh) against the
oracle, the golden files and size-independent properties (GPU)."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from helpers import GOLDEN, Oracle
from swarm_b200 import DerepResult, HostDb

CASES = [("derep_mix", False), ("derep_mix_z", True)]


def _check_texts(res, name):
    assert res.swarms_text() == (GOLDEN / f"{name}.d0.o").read_bytes()
    assert res.swarms_text(mothur=True) == (GOLDEN / f"{name}.d0.r.o").read_bytes()
    assert res.stats_text() == (GOLDEN / f"{name}.d0.s").read_bytes()
    assert res.structure_text() == (GOLDEN / f"{name}.d0.i").read_bytes()
    assert res.seeds_text() == (GOLDEN / f"{name}.d0.w").read_bytes()
    assert res.uclust_text() == (GOLDEN / f"{name}.d0.u").read_bytes()
    log = (GOLDEN / f"{name}.d0.log").read_text()
    assert log == f"\nNumber of swarms:  {res.clusters}\nLargest swarm:     {res.largest}\nHeaviest swarm:    {res.heaviest}\n"


@pytest.mark.parametrize("name,usearch", CASES)
def test_oracle_and_writers_match_reference(built, name, usearch):
    db = HostDb(GOLDEN / f"{name}.fasta", usearch_abundance=usearch)
    orc = Oracle(db)
    rep, mass, size, singles = orc.derep()
    # the oracle's chains are the clusters' members in index order
    for s in orc.d0_seeds[:50]:
        chain, a = [int(s)], int(orc.d0_next[s])
        while a:
            chain.append(a)
            a = int(orc.d0_next[a])
        assert chain == np.flatnonzero(rep == s).tolist()
    _check_texts(DerepResult(db, rep, mass, size, singles), name)
    # the writers split the clusters over several workers for large results: same bytes with the grain lowered to these inputs
    from swarm_b200.ffi import host_lib
    L = host_lib()
    try:
        L.swbh_set_writer_grain(1)
        for workers in (2, 5):
            L.swbh_set_threads(workers)
            _check_texts(DerepResult(db, rep, mass, size, singles), name)
    finally:
        L.swbh_set_writer_grain(200000)
        L.swbh_set_threads(0)


def test_length_is_part_of_the_key(built):
    # "A", "AA", "AAA" pack to the same zero words (A = 0): only the length tells them apart
    db = HostDb(text=b">a_3\nA\n>b_2\nAA\n>c_2\nAAA\n>d_1\naa\n>e_1\nA\n")
    rep, mass, size, singles = Oracle(db).derep()
    heads = [db.header(i) for i in range(db.n)]
    groups = sorted(sorted(heads[j] for j in np.flatnonzero(rep == r)) for r in np.unique(rep))
    assert groups == [["a_3", "e_1"], ["b_2", "d_1"], ["c_2"]]
    assert DerepResult(db, rep, mass, size, singles).swarms_text() == b"a_3 e_1\nb_2 d_1\nc_2\n"


def test_assemble_rejects_inconsistent_arrays(built):
    db = HostDb(text=b">a_3\nA\n>b_2\nAA\n")
    z32, z64 = np.zeros(2, np.uint32), np.zeros(2, np.uint64)
    with pytest.raises(ValueError):
        DerepResult(db, np.array([1, 1], np.uint32), z64, z32, z32)          # rep must point at a first occurrence
    with pytest.raises(ValueError):
        DerepResult(db, np.array([0, 1], np.uint32), z64, z32, z32)          # sizes must add up to n


def test_cli_empty_input_matches_reference(built, tmp_path):
    if not helpers.have_ref():
        pytest.skip("reference binary not built")
    fa = tmp_path / "empty.fa"
    fa.write_bytes(b"")
    cli = str(helpers.ROOT / "bin" / "swarm_b200")
    for flags in (["-d", "0"], ["-d", "0", "-r"], [], ["-r"], ["-d", "2"], ["-d", "2", "-r"]):
        outs = []
        for exe in (cli, str(helpers.REF_BIN)):
            o, l = tmp_path / "o", tmp_path / "l"
            p = subprocess.run([exe, *flags, "-o", str(o), "-l", str(l), str(fa)], capture_output=True)
            assert p.returncode == 0
            log = l.read_bytes()
            outs.append((o.read_bytes(), log[log.index(b"\nNumber of swarms"):]))
        assert outs[0] == outs[1], flags


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,usearch", CASES)
def test_engine_matches_oracle_and_reference(built, name, usearch):
    from swarm_b200 import Engine
    db = HostDb(GOLDEN / f"{name}.fasta", usearch_abundance=usearch)
    eng = Engine(0)
    eng.load(db)
    rep, mass, size, singles, k = eng.d0_dereplicate()
    want = Oracle(db).derep()
    for got, exp in zip((rep, mass, size, singles), want):
        assert np.array_equal(got, exp)
    res = DerepResult(db, rep, mass, size, singles)
    assert k == res.clusters
    _check_texts(res, name)


@pytest.mark.gpu
def test_engine_tiny_and_length_edge_cases(built):
    from swarm_b200 import Engine
    for text in (b">a_3\nA\n>b_2\nAA\n>c_2\nAAA\n>d_1\naa\n>e_1\nA\n", b">only_7\nACGTACGT\n",
                 b"".join(b">s%d_1\nACGTTGCA\n" % i for i in range(1000))):
        db = HostDb(text=text)
        eng = Engine(0)
        eng.load(db)
        rep, mass, size, singles, k = eng.d0_dereplicate()
        for got, exp in zip((rep, mass, size, singles), Oracle(db).derep()):
            assert np.array_equal(got, exp)
        assert k == len(np.unique(rep))


@pytest.mark.gpu
@pytest.mark.parametrize("n,copies,L", [(200_000, 5, 150), (50_000, 40, 64), (300_000, 1, 33)])
def test_engine_seeded_reads_properties(built, tmp_path, n, copies, L):
    """n unique sequences, each present 1..copies times: rep is the first occurrence, sums add up, and the result
    equals the oracle's (the oracle handles these sizes in seconds)"""
    from swarm_b200 import Engine
    fa = tmp_path / "u.fa"
    helpers.make_fasta(fa, n, L, 11)
    seqs = [l for l in fa.read_text().splitlines() if not l.startswith(">")]
    rng = np.random.default_rng(5)
    reps = rng.integers(1, copies + 1, size=n)
    idx = np.repeat(np.arange(n), reps)
    rng.shuffle(idx)
    ab = rng.choice([1, 1, 2, 7], size=len(idx))
    text = "".join(f">r{j}_{ab[j]}\n{seqs[i]}\n" for j, i in enumerate(idx)).encode()
    db = HostDb(text=text)
    eng = Engine(0)
    eng.load(db)
    rep, mass, size, singles, k = eng.d0_dereplicate()
    assert k == n and int(size.sum()) == db.n and int(mass.sum()) == int(db.abundance.sum())
    assert np.array_equal(rep[rep], rep) and np.all(rep <= np.arange(db.n))
    for got, exp in zip((rep, mass, size, singles), Oracle(db).derep()):
        assert np.array_equal(got, exp)
    # idempotence: dereplicating the representatives finds nothing to merge
    keep = np.flatnonzero(rep == np.arange(db.n))
    text2 = "".join(f">{db.header(int(i))}\n{seqs_of(db, int(i))}\n" for i in keep[:20000]).encode()
    db2 = HostDb(text=text2)
    eng2 = Engine(0)
    eng2.load(db2)
    rep2 = eng2.d0_dereplicate()[0]
    assert np.array_equal(rep2, np.arange(db2.n))


def seqs_of(db, i):
    w = db.words.reshape(db.n, db.stride)[i]
    return "".join("ACGT"[(int(w[p >> 5]) >> ((p & 31) * 2)) & 3] for p in range(int(db.len[i])))


@pytest.mark.gpu
def test_long_sequences(built):
    """sequences beyond 5 000 nt (the reference accepts up to 67 108 861 nt, src/db.cc:439-442): dereplicated (d = 0) and
    clustered at d = 1 by the JOIN path, whose global-memory multimap has no length limit; only the enumeration kernels
    (per-position tables on chip) and the d > 1 aligner refuse them"""
    from swarm_b200 import ENUM_HALF, ENUM_JOIN, Engine, EngineError
    rng = np.random.default_rng(3)
    a = "".join(rng.choice(list("ACGT"), 6000))
    b = a[:5999] + ("A" if a[5999] != "A" else "C")
    db = HostDb(text=f">x_3\n{a}\n>y_2\n{b}\n>z_1\n{a}\n>w_1\n{a[:5999]}\n".encode())
    eng = Engine(0)
    eng.load(db)
    rep, mass, size, singles, k = eng.d0_dereplicate()
    for got, exp in zip((rep, mass, size, singles), Oracle(db).derep()):
        assert np.array_equal(got, exp)
    assert k == 3 and rep.tolist() == [0, 1, 2, 0]          # order: x_3, y_2, w_1, z_1 (abundance, then header)
    eng.d1_index()
    with pytest.raises(EngineError) as e:                   # x and z are identical: the reference's duplicate fatal
        eng.d1_network()
    assert e.value.status == 3
    with pytest.raises(EngineError):
        eng.dn_cluster(2)
    eng.close()
    # a swarm of 9 000-nt amplicons: centroid, one-edit variants (substitution / deletion / insertion anywhere), two-edit ones
    cen = list(rng.choice(list("ACGT"), 9000))
    recs, seen = [("c_100", "".join(cen))], {"".join(cen)}
    for i in range(60):
        s = list(cen)
        for _ in range(1 + (i % 3 == 2)):
            p, r = int(rng.integers(0, len(s))), rng.random()
            if r < 0.6:
                s[p] = "ACGT"[("ACGT".index(s[p]) + int(rng.integers(1, 4))) % 4]
            elif r < 0.8:
                del s[p]
            else:
                s.insert(p, "ACGT"[int(rng.integers(0, 4))])
        s = "".join(s)
        if s not in seen:
            seen.add(s)
            recs.append((f"v{i}_{1 + i % 7}", s))
    db2 = HostDb(text="".join(f">{h}\n{s}\n" for h, s in recs).encode())
    orc = Oracle(db2)
    orc.network()
    orc.cluster()
    eng = Engine(0, enum_mode=ENUM_JOIN, collect_stats=1)
    eng.load(db2)
    eng.d1_index()
    eng.d1_network()
    links = eng.d1_export_links()
    links = links[np.lexsort((links[:, 1], links[:, 0]))]
    sw, gen, par = eng.d1_cluster()
    eng.close()
    assert np.array_equal(links, orc.links()) and len(links) > 30
    assert np.array_equal(sw, orc.swarm_of) and np.array_equal(gen, orc.generation) and np.array_equal(par, orc.parent)
    eng = Engine(0, enum_mode=ENUM_HALF)
    eng.load(db2)
    with pytest.raises(EngineError) as e:
        eng.d1_index()
    assert e.value.status == 5
    eng.close()


@pytest.mark.gpu
def test_cli_d0_outputs(built, tmp_path):
    cli = str(helpers.ROOT / "bin" / "swarm_b200")
    for name, flags in (("derep_mix", []), ("derep_mix_z", ["-z"])):
        outs = {k: tmp_path / k for k in "osiwu"}
        cmd = [cli, "-d", "0", *flags, "-l", str(tmp_path / "log")]
        for k, f in outs.items():
            cmd += ["-" + k, str(f)]
        p = subprocess.run(cmd + [str(GOLDEN / f"{name}.fasta")], capture_output=True)
        assert p.returncode == 0, p.stderr
        for k, f in outs.items():
            assert f.read_bytes() == (GOLDEN / f"{name}.d0.{k}").read_bytes(), (name, k)
        log = (tmp_path / "log").read_bytes()
        assert log[log.index(b"\nNumber of swarms"):] == (GOLDEN / f"{name}.d0.log").read_bytes()
        p = subprocess.run([cli, "-d", "0", "-r", *flags, "-l", os.devnull, str(GOLDEN / f"{name}.fasta")], capture_output=True)
        assert p.returncode == 0 and p.stdout == (GOLDEN / f"{name}.d0.r.o").read_bytes()
