"""GPU parity tests (-m gpu): the CUDA engine, called through the C ABI (include/swarm_b200.h),
against the CPU oracle and the committed golden outputs of the reference binary.  Bit-exact."""
from pathlib import Path

import numpy as np
import pytest

import helpers
from helpers import GOLDEN, Oracle
from swarm_b200 import ENUM_FULL, ENUM_HALF, ENUM_JOIN, D1Result, DnResult, Engine, EngineError, HostDb, scoring
from swarm_b200.ffi import network_text

pytestmark = pytest.mark.gpu
CASES = ["handmade", "c1_1k_150", "tie_1500_60", "short_600_20", "w32_400", "w64_400", "w65_300"]


def bases_of(db, i):
    w = db.words[i * db.stride:(i + 1) * db.stride]
    L = int(db.len[i])
    p = np.arange(L)
    return ((w[p >> 5] >> ((p & 31).astype(np.uint64) << np.uint64(1))) & np.uint64(3)).astype(np.int64)


def cpu_variants(seq, Z):
    """the reference's canonical microvariant set (src/variants.cc:184-249) with hashes from table Z"""
    L = len(seq)

    def H(s):
        h = np.uint64(0)
        for p, b in enumerate(s):
            h ^= Z[p, b]
        return int(h)
    out = {}
    for p in range(L):
        for b in range(4):
            if b != seq[p]:
                s = list(seq); s[p] = b
                out[(0 << 30) | (b << 28) | p] = H(s)
    for p in range(L):
        if p == 0 or seq[p] != seq[p - 1]:
            out[(1 << 30) | p] = H(list(seq[:p]) + list(seq[p + 1:]))
    for p in range(L + 1):
        for b in range(4):
            if p == 0 or b != seq[p - 1]:
                out[(2 << 30) | (b << 28) | p] = H(list(seq[:p]) + [b] + list(seq[p:]))
    return out


@pytest.mark.parametrize("name", ["handmade", "short_600_20", "w32_400", "w65_300", "c1_1k_150"])
def test_variant_enumeration_matches_cpu(built, name):
    db = HostDb(GOLDEN / f"{name}.fasta")
    eng = Engine(0)
    eng.load(db)
    rng = np.random.default_rng(1)
    seeds = sorted(set([0, db.n - 1] + list(rng.integers(0, db.n, 6))))
    if name == "handmade":
        seeds = list(range(db.n))
    for s in seeds:
        h, c, Z = eng.debug_variants(s, ENUM_FULL)
        seq = bases_of(db, s)
        want = cpu_variants(seq, Z)
        got = dict(zip(c.tolist(), h.tolist()))
        assert len(got) == len(c), "duplicate variant codes"
        assert got == want, (name, s)
        h2, c2, _ = eng.debug_variants(s, ENUM_HALF)
        half = dict(zip(c2.tolist(), h2.tolist()))
        # HALF: deletions + substitutions along the tournament 0->1 0->2 1->2 1->3 2->3 3->0
        arcs = {(0, 1), (0, 2), (1, 2), (1, 3), (2, 3), (3, 0)}
        want_half = {k: v for k, v in want.items()
                     if (k >> 30) == 1 or ((k >> 30) == 0 and (int(seq[k & 0x0FFFFFFF]), (k >> 28) & 3) in arcs)}
        assert half == want_half, (name, s)
    eng.close()


def run_engine(db, mode, ncb=False, **opt):
    eng = Engine(0, enum_mode=mode, collect_stats=1, **opt)
    eng.load(db)
    eng.d1_index()
    eng.d1_network(no_cluster_breaking=ncb)
    links = eng.d1_export_links()
    links = links[np.lexsort((links[:, 1], links[:, 0]))]
    sw, gen, par = eng.d1_cluster()
    rp, col = eng.d1_get_network()
    stats = eng.stats()
    eng.close()
    return links, sw, gen, par, rp, col, stats


@pytest.mark.parametrize("mode", [ENUM_FULL, ENUM_HALF, ENUM_JOIN])
@pytest.mark.parametrize("name", CASES)
def test_golden_cases(built, name, mode):
    db = HostDb(GOLDEN / f"{name}.fasta")
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    links, sw, gen, par, rp, col, stats = run_engine(db, mode)
    assert np.array_equal(links, orc.links()), "directed link set differs from the oracle"
    assert np.array_equal(sw, orc.swarm_of)
    assert np.array_equal(gen, orc.generation)
    assert np.array_equal(par, orc.parent)
    res = D1Result(db, sw, gen, par)
    assert res.swarms_text() == (GOLDEN / f"{name}.o").read_bytes()
    assert res.stats_text() == (GOLDEN / f"{name}.s").read_bytes()
    assert res.structure_text() == (GOLDEN / f"{name}.i").read_bytes()
    assert network_text(db, rp, col) == (GOLDEN / f"{name}.j").read_bytes()
    if mode == ENUM_JOIN and db.len.min() >= 16:
        assert stats["variants"] == 2 * db.n             # two K-mer entries per amplicon, no enumeration
    if mode == ENUM_FULL:   # the full enumeration probes exactly the reference's variant count
        assert stats["variants"] == int(orc.net_stats[0])


@pytest.mark.parametrize("mode", [ENUM_FULL, ENUM_HALF, ENUM_JOIN])
@pytest.mark.parametrize("name", ["handmade", "tie_1500_60", "c1_1k_150"])
def test_no_cluster_breaking(built, name, mode):
    db = HostDb(GOLDEN / f"{name}.fasta")
    links, sw, gen, par, *_ = run_engine(db, mode, ncb=True)
    res = D1Result(db, sw, gen, par)
    assert res.swarms_text() == (GOLDEN / f"{name}.n.o").read_bytes()


@pytest.mark.parametrize("name", CASES)
def test_lean_half_kernel_matches_first_kernel(built, name):
    db = HostDb(GOLDEN / f"{name}.fasta")
    orc = Oracle(db)
    orc.network()
    l1, *_r1, s1 = run_engine(db, ENUM_HALF, net_kernel=1)
    l2, *_r2, s2 = run_engine(db, ENUM_HALF, net_kernel=2)
    assert np.array_equal(l1, l2) and np.array_equal(l2, orc.links())
    assert s1["variants"] == s2["variants"] and s1["filter_pass"] == s2["filter_pass"]


# JOIN flavours: tile store (default: 8-byte entries, rows gathered), the same with 8-record tile slots (most records take
# the overflow path of k_ts_big), both again with fat records (entry + packed row, the sharded-database layout), the global hash multimap of d1_join.cuh, and r1's count/scan/scatter tile join (normal and tiny tiles)
JOIN_FLAVOURS = [{}, {"tile_cmax": 8}, {"tile_rows": 1}, {"tile_rows": 1, "tile_cmax": 8}, {"join_kernel": 1}, {"join_kernel": 2},
                 {"join_kernel": 2, "tile_cmax": 8}]


@pytest.mark.parametrize("flavour", JOIN_FLAVOURS)
@pytest.mark.parametrize("name", CASES)
def test_join_flavours_golden(built, name, flavour):
    db = HostDb(GOLDEN / f"{name}.fasta")
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    for ncb in (False, True):
        links, sw, gen, par, rp, col, stats = run_engine(db, ENUM_JOIN, ncb=ncb, **flavour)
        if not ncb:
            assert np.array_equal(links, orc.links())
            assert np.array_equal(sw, orc.swarm_of) and np.array_equal(gen, orc.generation) and np.array_equal(par, orc.parent)
            assert network_text(db, rp, col) == (GOLDEN / f"{name}.j").read_bytes()
        elif (GOLDEN / f"{name}.n.o").exists():
            assert D1Result(db, sw, gen, par).swarms_text() == (GOLDEN / f"{name}.n.o").read_bytes()


@pytest.mark.parametrize("flavour", JOIN_FLAVOURS)
@pytest.mark.parametrize("n,L,seed,mode_ab", [(60000, 150, 42, 0), (40000, 80, 9, 1), (30000, 400, 5, 0), (20000, 31, 4, 1), (15000, 700, 6, 0)])
def test_join_flavours_seeded(built, tmp_path, n, L, seed, mode_ab, flavour):
    fa = helpers.make_fasta(tmp_path / "s.fa", n, L, seed, mode_ab)
    db = HostDb(fa)
    orc = Oracle(db)
    orc.network()
    links, *_ = run_engine(db, ENUM_JOIN, **flavour)
    assert np.array_equal(links, orc.links())


@pytest.mark.parametrize("group", [300, 3000])
def test_join_dense_group(built, tmp_path, group):
    """one dense group sharing both K-mers (every amplicon is a one- or two-edit variant of the same centroid, edits
    in the middle).  3000: the tile holding it overflows shared memory and goes through k_tile_join_big; 300: it
    fits, but its 45 k same-key pairs fill the tile's pair queue many times (resumable walk of k_tile_join)."""
    rng = np.random.default_rng(5)
    cen = rng.integers(0, 4, 200)
    seqs = {"".join("ACGT"[b] for b in cen)}
    while len(seqs) < group:
        s = cen.copy()
        p = int(rng.integers(70, 130))
        s[p] = (s[p] + int(rng.integers(1, 4))) % 4
        if rng.random() < 0.5:
            q = int(rng.integers(70, 130))
            s[q] = (s[q] + int(rng.integers(1, 4))) % 4
        seqs.add("".join("ACGT"[b] for b in s))
    text = "".join(f">d{i}_{1 + (i * 7919) % 50}\n{s}\n" for i, s in enumerate(sorted(seqs))).encode()
    db = HostDb(text=text)
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    for flavour in JOIN_FLAVOURS:
        links, sw, gen, par, *_ = run_engine(db, ENUM_JOIN, **flavour)
        assert np.array_equal(links, orc.links())
        assert np.array_equal(sw, orc.swarm_of) and np.array_equal(gen, orc.generation) and np.array_equal(par, orc.parent)


def test_skew_fallback_to_enumeration(built):
    """dense data: 9 000 amplicons share both K-mers, so the two tiles that hold them overflow by 8 k records each and
    the pairwise sweep would cost 8e7 pair tests: swb200_d1_network abandons the join for the linear HALF enumeration
    (the reference is linear in density, src/algod1.cc:606-670).  Same links, same swarms; with the fallback disabled the
    sweep itself must give them too."""
    rng = np.random.default_rng(11)
    cen = rng.integers(0, 4, 200)
    seqs = {"".join("ACGT"[b] for b in cen)}
    while len(seqs) < 9000:
        s = cen.copy()
        for p in rng.integers(66, 134, 2 if rng.random() < 0.97 else 1):
            s[p] = (s[p] + int(rng.integers(1, 4))) % 4
        seqs.add("".join("ACGT"[b] for b in s))
    text = "".join(f">d{i}_{1 + (i * 7919) % 50}\n{s}\n" for i, s in enumerate(sorted(seqs))).encode()
    db = HostDb(text=text)
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    for opt, fallbacks in (({}, 1), ({"skew_fallback": 0}, 0)):
        links, sw, gen, par, _rp, _col, stats = run_engine(db, ENUM_JOIN, **opt)
        assert stats["skew_fallbacks"] == fallbacks and stats["tile_overflow"] > 15000
        assert np.array_equal(links, orc.links())
        assert np.array_equal(sw, orc.swarm_of) and np.array_equal(gen, orc.generation) and np.array_equal(par, orc.parent)


def test_duplicates_rejected(built):
    db = HostDb(text=b">a_3\nACGTACGTACGTACGTACGTACGTACGTACGTACGTA\n>b_2\nACGTACGTACGTACGTACGTACGTACGTACGTACGTA\n>c_1\nACGTACGA\n")
    eng = Engine(0, enum_mode=ENUM_HALF)
    eng.load(db)
    with pytest.raises(EngineError) as e:
        eng.d1_index()
    assert e.value.status == 3
    eng.close()
    db = HostDb(text=b">a_3\nACGTACGTACGTACGTACGTACGTACGTACGTACGTA\n>b_2\nACGTACGTACGTACGTACGTACGTACGTACGTACGTA\n>c_1\nACGTACGAGGGGGGGGGTTTTT\n")
    eng = Engine(0, enum_mode=ENUM_JOIN)       # JOIN reports duplicates from the network phase (ed = 0 candidates)
    eng.load(db)
    eng.d1_index()
    with pytest.raises(EngineError) as e:
        eng.d1_network()
    assert e.value.status == 3
    eng.close()


@pytest.mark.parametrize("n,L,seed,mode_ab", [(60000, 150, 42, 0), (40000, 80, 9, 1), (30000, 400, 5, 0), (20000, 31, 4, 1)])
def test_seeded_sets_vs_oracle(built, tmp_path, n, L, seed, mode_ab):
    fa = helpers.make_fasta(tmp_path / "s.fa", n, L, seed, mode_ab)
    db = HostDb(fa)
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    for mode, opt in ((ENUM_FULL, {"cluster_kernel": 1}), (ENUM_HALF, {"net_kernel": 1, "cluster_kernel": 2}), (ENUM_HALF, {"net_kernel": 2, "cluster_kernel": 3}), (ENUM_JOIN, {}),
                      (ENUM_JOIN, {"cluster_kernel": 4, "tile_rows": 1}),
                      (ENUM_JOIN, {"cluster_kernel": 6})):
        links, sw, gen, par, *_ = run_engine(db, mode, **opt)
        assert np.array_equal(links, orc.links())
        assert np.array_equal(sw, orc.swarm_of)
        assert np.array_equal(gen, orc.generation)
        assert np.array_equal(par, orc.parent)


@pytest.mark.parametrize("cluster_kernel", [0, 6])
def test_packed_relaxation_word_and_its_fallback(built, tmp_path, cluster_kernel):
    """the clustering kernels relax ONE word swarm | generation | parent (cluster_pack, default): same arrays as r1's key + parent
    pass (cluster_pack=0) and as the oracle; a swarm deeper than the generation field holds (forced: 2 generation bits) is
    detected and redone with 32-bit generations, transparently."""
    fa = helpers.make_fasta(tmp_path / "s.fa", 50000, 120, 77, 1)           # tie-heavy: links in both directions
    db = HostDb(fa)
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    assert int(orc.generation.max()) >= 3
    for opt, reruns in (({}, 0), ({"cluster_pack": 0}, 0), ({"cluster_gen_bits": 2}, 1), ({"cluster_gen_bits": 31}, 0)):
        eng = Engine(0, cluster_kernel=cluster_kernel, **opt)
        eng.load(db)
        eng.d1_index()
        eng.d1_network()
        sw, gen, par = eng.d1_cluster()
        st = eng.stats()
        eng.close()
        assert st["cluster_unpacked_reruns"] == reruns, (opt, st)
        assert np.array_equal(sw, orc.swarm_of) and np.array_equal(gen, orc.generation) and np.array_equal(par, orc.parent), opt


def test_filter_sizes_and_sharding_give_identical_links(built, tmp_path):
    fa = helpers.make_fasta(tmp_path / "s.fa", 50000, 150, 77, 0)
    db = HostDb(fa)
    base, *_ = run_engine(db, ENUM_HALF)
    for bps in (2, 4):
        links, *_ = run_engine(db, ENUM_HALF, bloom_bytes_per_slot=bps)
        assert np.array_equal(links, base)
    for mode in (ENUM_HALF, ENUM_JOIN):      # seed shards (enumeration) / hash-range shards of the K-mer table (join)
        parts = []
        for r in range(3):
            links, *_ = run_engine(db, mode, shard_rank=r, shard_world=3)
            parts.append(links)
        allp = np.concatenate(parts)
        allp = allp[np.lexsort((allp[:, 1], allp[:, 0]))]
        assert np.array_equal(allp, base)


def test_large_set_properties(built, tmp_path):
    """size-independent checks at a size the oracle would take minutes for: FULL == HALF, clustering is
    a fixed point (every link stays inside one swarm or points from a smaller label), seeds are minima."""
    fa = helpers.make_fasta(tmp_path / "big.fa", 1000000, 150, 11, 0)
    db = HostDb(fa)
    lf, swf, genf, parf, *_ = run_engine(db, ENUM_FULL, cluster_kernel=1)
    lh, swh, genh, parh, *_ = run_engine(db, ENUM_HALF, net_kernel=2)
    l1, *_ = run_engine(db, ENUM_HALF, net_kernel=1)
    assert np.array_equal(l1, lh)
    lj, swj, genj, parj, *_ = run_engine(db, ENUM_JOIN)
    assert np.array_equal(lj, lh) and np.array_equal(swj, swh)
    assert np.array_equal(lf, lh) and np.array_equal(swf, swh) and np.array_equal(genf, genh) and np.array_equal(parf, parh)
    src, dst = lf[:, 0].astype(np.int64), lf[:, 1].astype(np.int64)
    assert np.all(db.abundance[src] >= db.abundance[dst])
    assert np.all(swf[dst] <= swf[src])              # min-label fixed point over directed links
    assert np.all(swf <= np.arange(db.n))            # a seed is the smallest id of its swarm
    roots = swf == np.arange(db.n)
    assert np.all(genf[roots] == 0) and np.all(parf[roots] == 0xFFFFFFFF)
    nz = ~roots
    assert np.all(genf[parf[nz].astype(np.int64)] + 1 == genf[nz]) and np.all(swf[parf[nz].astype(np.int64)] == swf[nz])
    if helpers.with_ref():
        r = helpers.run_ref(fa, outputs=("o",), threads=8)
        res = D1Result(db, swf, genf, parf)
        assert res.swarms_text() == r["o"]


@pytest.mark.parametrize("fast_kernel", [1, 2])
@pytest.mark.parametrize("boundary,suffix", [(3, "f"), (10, "f.b10")])
@pytest.mark.parametrize("name", CASES)
def test_fastidious_golden(built, name, boundary, suffix, fast_kernel):
    db = HostDb(GOLDEN / f"{name}.fasta")
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    grafts = orc.fastidious(boundary=boundary)
    eng = Engine(0, fast_kernel=fast_kernel)
    eng.load(db)
    eng.d1_index()
    eng.d1_network()
    sw, gen, par = eng.d1_cluster()
    gc, nl, nh = eng.d1_fastidious(boundary=boundary)
    eng.close()
    if grafts >= 0:
        assert np.array_equal(gc, orc.graft_raw), "graft candidates differ from the oracle"
    else:
        assert np.all(gc == 0xFFFFFFFF)
    res = D1Result(db, sw, gen, par, graft_cand=gc, boundary=boundary)
    assert res.swarms_text() == (GOLDEN / f"{name}.{suffix}.o").read_bytes()
    if suffix == "f":
        assert res.stats_text() == (GOLDEN / f"{name}.f.s").read_bytes()
        assert res.structure_text() == (GOLDEN / f"{name}.f.i").read_bytes()


@pytest.mark.parametrize("fast_kernel", [1, 2])
@pytest.mark.parametrize("n,L,seed,mode_ab,boundary", [(30000, 150, 21, 0, 3), (20000, 60, 22, 1, 4), (8000, 400, 23, 0, 3)])
def test_fastidious_seeded_vs_oracle(built, tmp_path, n, L, seed, mode_ab, boundary, fast_kernel):
    fa = helpers.make_fasta(tmp_path / "s.fa", n, L, seed, mode_ab)
    db = HostDb(fa)
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    grafts = orc.fastidious(boundary=boundary)
    eng = Engine(0, fast_kernel=fast_kernel)
    eng.load(db)
    eng.d1_index()
    eng.d1_network()
    sw, gen, par = eng.d1_cluster()
    gc, nl, nh = eng.d1_fastidious(boundary=boundary)
    st = eng.stats()
    eng.close()
    assert grafts > 0 and np.array_equal(gc, orc.graft_raw)
    if fast_kernel == 1:
        assert st["fast_light_variants"] == int(orc.fast_stats[0])      # same microvariant count as the reference's light pass
        assert st["fast_heavy_variants"] == int(orc.fast_stats[1])
    else:
        assert st["fast_light_variants"] == 3 * nl                      # join: three K-mer entries per light amplicon
    res_g = D1Result(db, sw, gen, par, graft_cand=gc, boundary=boundary)
    res_o = D1Result(db, orc.swarm_of, orc.generation, orc.parent, graft_cand=orc.graft_cand, boundary=boundary)
    assert res_g.swarms_text() == res_o.swarms_text() and res_g.grafts == grafts
    if helpers.with_ref():
        r = helpers.run_ref(fa, "-f", "-b", str(boundary), outputs=("o", "s", "i"), threads=1)
        assert res_g.swarms_text() == r["o"] and res_g.stats_text() == r["s"] and res_g.structure_text() == r["i"]


DN_CASES = [("handmade", 2, False, None, "d2"), ("c1_1k_150", 2, False, None, "d2"), ("short_600_20", 2, False, None, "d2"),
            ("short_600_20", 3, False, None, "d3"), ("tie_1500_60", 2, False, None, "d2"), ("tie_1500_60", 2, True, None, "d2n"),
            ("w65_300", 3, False, None, "d3"), ("w64_400", 4, False, None, "d4"), ("w32_400", 2, False, (3, 2, 5, 3), "d2pen"),
            ("l400_250", 2, False, None, "d2"),
            # band half-width > 15 (d >= 7 with the default scoring): k_dn_align_wide; the reference runs its 16-bit aligner
            ("c1_1k_150", 7, False, None, "d7"), ("w65_300", 9, False, None, "d9"), ("w64_400", 12, False, None, "d12"),
            ("l400_250", 7, False, None, "d7"), ("handmade", 8, False, None, "d8")]


@pytest.mark.parametrize("dn_filter", [0, 1])
@pytest.mark.parametrize("name,d,ncb,pen,tag", DN_CASES)
def test_dn_golden(built, name, d, ncb, pen, tag, dn_filter):
    db = HostDb(GOLDEN / f"{name}.fasta", check_dup_sequences=True)
    p = scoring(*pen) if pen else scoring()
    orc = Oracle(db)
    osw, ogen, opar, opd = orc.dn_cluster(d, no_cluster_breaking=ncb, pen=p)
    eng = Engine(0, dn_filter=dn_filter)
    eng.load(db)
    sw, gen, par, pd = eng.dn_cluster(d, no_cluster_breaking=ncb, penalties=p)
    eng.close()
    assert np.array_equal(sw, osw) and np.array_equal(gen, ogen) and np.array_equal(par, opar) and np.array_equal(pd, opd)
    res = DnResult(db, sw, gen, par, pd)
    assert res.swarms_text() == (GOLDEN / f"{name}.{tag}.o").read_bytes()
    assert res.stats_text() == (GOLDEN / f"{name}.{tag}.s").read_bytes()
    assert res.structure_text() == (GOLDEN / f"{name}.{tag}.i").read_bytes()


@pytest.mark.parametrize("dn_filter", [0, 1])
@pytest.mark.parametrize("n,L,seed,mode_ab,d", [(4000, 100, 31, 0, 2), (3000, 60, 32, 1, 3), (2500, 400, 33, 0, 2), (2000, 150, 34, 1, 5),
                                                (1500, 120, 35, 0, 8), (1200, 90, 36, 1, 10), (300, 70, 37, 0, 255)])
def test_dn_seeded_vs_oracle_and_reference(built, tmp_path, n, L, seed, mode_ab, d, dn_filter):
    fa = helpers.make_fasta(tmp_path / "s.fa", n, L, seed, mode_ab)
    db = HostDb(fa, check_dup_sequences=True)
    orc = Oracle(db)
    osw, ogen, opar, opd = orc.dn_cluster(d)
    eng = Engine(0, dn_filter=dn_filter)
    eng.load(db)
    sw, gen, par, pd = eng.dn_cluster(d)
    st = eng.stats()
    eng.close()
    assert np.array_equal(sw, osw) and np.array_equal(gen, ogen) and np.array_equal(par, opar) and np.array_equal(pd, opd)
    assert st["dn_links"] >= int(orc.dn_stats[2])      # all directed links vs the greedy loop's accepted ones
    if helpers.with_ref():
        r = helpers.run_ref(fa, "-d", str(d), outputs=("o", "s", "i"), threads=4)
        res = DnResult(db, sw, gen, par, pd)
        assert res.swarms_text() == r["o"]
        # Unrelated sequences (random pairs ~45 differences apart, only linked at d = 255) have many co-optimal
        # alignments, and the reference's SIMD aligner (src/search16.cc) and its own scalar one (src/nw.cc, which the
        # oracle and the engine follow) then report different difference counts for ~2 % of such pairs; on related
        # sequences they agree at any distance (test_dn_wide_related_sequences).  So the per-link columns are compared
        # where links join related sequences.
        if d <= 30:
            assert res.stats_text() == r["s"] and res.structure_text() == r["i"]


@pytest.mark.parametrize("d", [20, 60])
def test_dn_wide_related_sequences(built, tmp_path, d):
    """one seed + 600 variants 3..60 edits away, clustered at a d far beyond the register band (full-matrix
    k_dn_align_wide): engine == oracle == reference, including every link's difference count"""
    fa = helpers.make_variant_fasta(tmp_path / "v.fa", 600, 150, 5)
    db = HostDb(fa, check_dup_sequences=True)
    orc = Oracle(db)
    osw, ogen, opar, opd = orc.dn_cluster(d)
    eng = Engine(0)
    eng.load(db)
    sw, gen, par, pd = eng.dn_cluster(d)
    eng.close()
    assert np.array_equal(sw, osw) and np.array_equal(gen, ogen) and np.array_equal(par, opar) and np.array_equal(pd, opd)
    if helpers.with_ref():
        r = helpers.run_ref(fa, "-d", str(d), outputs=("o", "s", "i"), threads=4)
        res = DnResult(db, sw, gen, par, pd)
        assert res.swarms_text() == r["o"] and res.stats_text() == r["s"] and res.structure_text() == r["i"]


def _engine_vs_oracle(db, **opt):
    orc = Oracle(db)
    assert orc.network() is not None
    orc.cluster()
    links, sw, gen, par, *_ = run_engine(db, opt.pop("mode", ENUM_JOIN), **opt)
    assert np.array_equal(links, orc.links())
    assert np.array_equal(sw, orc.swarm_of) and np.array_equal(gen, orc.generation) and np.array_equal(par, orc.parent)


def test_edge_cases(built):
    # a single amplicon; two unrelated singletons; a pure chain of equal abundances; one-nucleotide sequences
    for text in (b">a_1\nACGTACGTACGTACGTACGTAGCTAGCTAGGATC\n",
                 b">a_2\nACGTACGTACGTACGTACGTAGCTAGCTAGGATC\n>b_1\nTTTTTTTTTTTTTTTTTTTTTTGGGGGGGGGGGGG\n",
                 b">a_1\nAAAAAAAAAAAAAAAAAAAA\n>b_1\nAAAAAAAAAAAAAAAAAAAC\n>c_1\nAAAAAAAAAAAAAAAAAACC\n>d_1\nAAAAAAAAAAAAAAAAACCC\n>e_1\nAAAAAAAAAAAAAAAACCCC\n",
                 b">a_5\nA\n>b_4\nC\n>c_3\nAC\n>d_2\nACG\n>e_1\nG\n"):
        db = HostDb(text=text)
        for mode in (ENUM_FULL, ENUM_HALF, ENUM_JOIN):
            _engine_vs_oracle(db, mode=mode)


def test_long_sequences_use_the_general_kernels(built, tmp_path):
    # > 990 nt: the lean HALF kernel and its fused table do not apply; JOIN pieces are capped at 64 nt
    fa = helpers.make_fasta(tmp_path / "long.fa", 1500, 1100, 3, 0)
    db = HostDb(fa)
    assert db.longest > 1000
    for mode in (ENUM_HALF, ENUM_JOIN, ENUM_FULL):
        _engine_vs_oracle(db, mode=mode)
    orc = Oracle(db); orc.network(); orc.cluster()
    if orc.fastidious(boundary=3) >= 0:
        for fk in (1, 2):
            eng = Engine(0, fast_kernel=fk)
            eng.load(db); eng.d1_index(); eng.d1_network(); eng.d1_cluster()
            gc, *_ = eng.d1_fastidious(boundary=3)
            eng.close()
            assert np.array_equal(gc, orc.graft_raw)


def test_mixed_lengths(built, tmp_path):
    # lengths from 30 to 300 in one database: piece length K is set by the shortest sequence
    parts = []
    for i, L in enumerate((30, 75, 150, 300)):
        p = helpers.make_fasta(tmp_path / f"p{i}.fa", 800, L, 50 + i, i % 2)
        parts.append(open(p, "rb").read().replace(b">s", b">l%d_" % L))
    db = HostDb(text=b"".join(parts))
    for mode in (ENUM_HALF, ENUM_JOIN):
        _engine_vs_oracle(db, mode=mode)
    orc = Oracle(db); orc.network(); orc.cluster(); g = orc.fastidious(boundary=3)
    eng = Engine(0)
    eng.load(db); eng.d1_index(); eng.d1_network(); eng.d1_cluster()
    gc, *_ = eng.d1_fastidious(boundary=3)
    eng.close()
    if g >= 0:
        assert np.array_equal(gc, orc.graft_raw)


def test_dist_clustering_world1(built, tmp_path):
    """swb200_d1_cluster_dist with a single rank (every link is local): same arrays as swb200_d1_cluster; twice, to
    check that the inboxes and barrier epochs are reusable."""
    import torch
    from swarm_b200.ffi import dist_buffer_bytes
    fa = helpers.make_fasta(tmp_path / "s.fa", 50000, 150, 17, 0)
    db = HostDb(fa)
    eng = Engine(0)
    eng.load(db)
    eng.d1_index()
    eng.d1_network()
    sw, gen, par = eng.d1_cluster()
    nbytes = dist_buffer_bytes(db.n, 1)
    buf = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device="cuda")
    eng.dist_setup(0, 1, [buf.data_ptr()], nbytes)
    for _ in range(2):
        out = {k: np.empty(db.n, dtype=np.uint32) for k in ("swarm_of", "generation", "parent")}
        eng.d1_cluster_dist(out)
        assert np.array_equal(out["swarm_of"], sw) and np.array_equal(out["generation"], gen) and np.array_equal(out["parent"], par)
    eng.close()


def test_dist_clustering_world2(built, tmp_path):
    """two ranks on two GPUs (torchrun): sharded upload, sharded join, peer-memory clustering == single-GPU rows"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fa = helpers.make_fasta(tmp_path / "s.fa", 300000, 150, 23, 0)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(Path(__file__).resolve().parent / "dist_worker.py"), str(fa)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("dist ok") == 2


@pytest.mark.parametrize("name", ["handmade", "tie_1500_60", "c1_1k_150"])
def test_compact_loader_gives_the_same_database(built, name):
    """swb200_load_db_compact (u16 lengths + abundance runs, expanded on the device) == swb200_load_db: same links,
    same clustering, same grafts (the fastidious pass reads the abundances again for the swarm masses)"""
    from swarm_b200.ffi import compact_form
    db = HostDb(GOLDEN / f"{name}.fasta")
    outs = []
    for compact in (False, True):
        eng = Engine(0)
        if compact:
            l16, rab, rst = compact_form(db.len, db.abundance)
            assert rst[0] == 0 and rst[-1] == db.n and len(rab) == len(np.unique(db.abundance))
            eng.load_db_compact(np.ascontiguousarray(db.words), db.stride, l16, rab, rst)
        else:
            eng.load(db)
        eng.d1_index()
        eng.d1_network()
        links = eng.d1_export_links()
        links = links[np.lexsort((links[:, 1], links[:, 0]))]
        sw, gen, par = eng.d1_cluster()
        gc, nl, nh = eng.d1_fastidious(boundary=3)
        outs.append((links, sw, gen, par, gc, nl, nh))
        eng.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    res = D1Result(db, outs[1][1], outs[1][2], outs[1][3], graft_cand=outs[1][4] if outs[1][5] and outs[1][6] else None)
    assert res.swarms_text() == (GOLDEN / f"{name}.f.o").read_bytes()


def test_shard_compact_loader_world1(built, tmp_path):
    """swb200_load_db_shard_compact with a single shard (no exchange needed) + swb200_db_commit == swb200_load_db: same links, same
    swarms; the abundance array is rebuilt on the device from the run table alone.  (Two shards + NCCL all-gather: tests/dist_worker.py,
    bench.py --gpus N.)"""
    from swarm_b200.ffi import compact_form
    fa = helpers.make_fasta(tmp_path / "s.fa", 30000, 150, 5, 1)
    db = HostDb(fa)
    l16, rab, rst = compact_form(db.len, db.abundance)
    ref = run_engine(db, ENUM_JOIN)
    eng = Engine(0)
    eng.load_db_shard_compact(db.words, db.stride, l16, db.n, 0, rab, rst)
    eng.db_commit()
    eng.d1_index()
    eng.d1_network()
    links = eng.d1_export_links()
    links = links[np.lexsort((links[:, 1], links[:, 0]))]
    sw, gen, par = eng.d1_cluster()
    eng.close()
    assert np.array_equal(links, ref[0]) and np.array_equal(sw, ref[1]) and np.array_equal(gen, ref[2]) and np.array_equal(par, ref[3])
    eng = Engine(0)
    bad = rst.copy()
    bad[-1] -= 1                                                 # the runs do not cover [0, n)
    with pytest.raises(EngineError):
        eng.load_db_shard_compact(db.words, db.stride, l16, db.n, 0, rab, bad)
    eng.close()


def test_compact_loader_rejects_bad_runs(built):
    db = HostDb(GOLDEN / "handmade.fasta")
    eng = Engine(0)
    l16 = db.len.astype(np.uint16)
    w = np.ascontiguousarray(db.words)
    for rst in ([1, db.n], [0, db.n - 1], [0, 5, 5, db.n], [0, 7, 3, db.n]):
        rst = np.array(rst, dtype=np.uint32)
        with pytest.raises(EngineError):
            eng.load_db_compact(w, db.stride, l16, np.ones(len(rst) - 1, dtype=np.uint64), rst)
    eng.close()
