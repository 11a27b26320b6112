"""Test-side bindings: the CPU oracle (oracle/liboracle.so), the reference binary (oracle/_ref/swarm,
when present), the synthetic generator, and output canonicalisation.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
REF_BIN = ROOT / "oracle" / "_ref" / "swarm"
NONE = 0xFFFFFFFF

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def _p(a, ty):
    return a.ctypes.data_as(ty)


class OrcDb(C.Structure):
    _fields_ = [("n", C.c_uint32), ("longest", C.c_uint32), ("words", _u64p), ("off", _u64p),
                ("len", _u32p), ("abundance", _u64p)]


class OrcVar(C.Structure):
    _fields_ = [("hash", C.c_uint64), ("pos", C.c_uint32), ("type", C.c_uint8), ("base", C.c_uint8), ("pad", C.c_uint16)]


_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        L = C.CDLL(str(ROOT / "oracle" / "liboracle.so"))
        L.orc_zobrist_init.argtypes = [C.c_uint32]
        L.orc_zobrist_value.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_zobrist_value.restype = C.c_uint64
        L.orc_zobrist_hash.argtypes = [_u64p, C.c_uint32]
        L.orc_zobrist_hash.restype = C.c_uint64
        L.orc_mt19937_64_next.restype = C.c_uint64
        L.orc_generate_variants.argtypes = [_u64p, C.c_uint32, C.c_uint64, C.POINTER(OrcVar)]
        L.orc_generate_variants.restype = C.c_uint32
        L.orc_hashtable_size.argtypes = [C.c_uint64]
        L.orc_hashtable_size.restype = C.c_uint64
        L.orc_d1_network.argtypes = [C.POINTER(OrcDb), C.c_int, _u32p, _u32p, C.POINTER(_u32p), _u64p, _u64p]
        L.orc_d1_cluster.argtypes = [C.POINTER(OrcDb), _u32p, _u32p, _u32p] + [_u32p] * 4 + [_u32p] * 5 + [_u64p] * 2
        L.orc_d1_cluster.restype = C.c_uint32
        L.orc_d1_fastidious.argtypes = [C.POINTER(OrcDb), C.c_uint64, C.c_uint32, C.c_uint32, _u32p, _u32p,
                                        _u32p, _u32p, _u32p, _u32p, _u64p, _u64p, _u8p, _u32p, _u32p, _u64p]
        L.orc_d1_fastidious.restype = C.c_int64
        L.orc_findqgrams.argtypes = [_u64p, C.c_uint32, _u8p]
        L.orc_qgram_diff.argtypes = [_u8p, _u8p]
        L.orc_qgram_diff.restype = C.c_uint64
        L.orc_scoring.argtypes = [C.c_int64] * 4 + [C.POINTER(C.c_int64)]
        L.orc_nw_diffs.argtypes = [_u64p, C.c_uint32, _u64p, C.c_uint32, C.c_int64, C.c_int64, C.c_int64, _u64p]
        L.orc_nw_diffs.restype = C.c_uint64
        L.orc_dn_cluster.argtypes = [C.POINTER(OrcDb), C.c_uint32, C.c_int, C.POINTER(C.c_int64)] + [_u32p] * 6 + [_u64p]
        L.orc_dn_cluster.restype = C.c_uint32
        L.orc_d0_dereplicate.argtypes = [C.POINTER(OrcDb), _u32p, _u32p, _u32p, _u64p, _u32p, _u32p]
        L.orc_d0_dereplicate.restype = C.c_uint32
        L.orc_free.argtypes = [C.c_void_p]
        _orc = L
    return _orc


class Oracle:
    """CPU restatement of the d=1 path over a HostDb (fixed-stride words)."""

    def __init__(self, db):
        self.db = db
        n = db.n
        self.words = np.ascontiguousarray(db.words, dtype=np.uint64)
        self.off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(db.stride))
        self.len = np.ascontiguousarray(db.len, dtype=np.uint32)
        self.ab = np.ascontiguousarray(db.abundance, dtype=np.uint64)
        self.c = OrcDb(n, db.longest, _p(self.words, _u64p), _p(self.off, _u64p), _p(self.len, _u32p), _p(self.ab, _u64p))

    def network(self, no_cluster_breaking=False):
        L = oracle_lib()
        n = self.db.n
        ls = np.zeros(n, dtype=np.uint32)
        lc = np.zeros(n, dtype=np.uint32)
        net = _u32p()
        m = C.c_uint64()
        st = np.zeros(4, dtype=np.uint64)
        rc = L.orc_d1_network(C.byref(self.c), int(no_cluster_breaking), _p(ls, _u32p), _p(lc, _u32p), C.byref(net),
                              C.byref(m), _p(st, _u64p))
        if rc != 0:
            return None
        arr = np.ctypeslib.as_array(net, shape=(max(m.value, 1),))[: m.value].copy()
        L.orc_free(net)
        self.link_start, self.link_count, self.net, self.net_stats = ls, lc, arr, st
        return ls, lc, arr

    def links(self):
        """directed links as a sorted (src, dst) array"""
        src = np.repeat(np.arange(self.db.n, dtype=np.uint32), self.link_count)
        # rows are contiguous and in seed order in the oracle's network
        pairs = np.stack([src, self.net], axis=1)
        return pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]

    def cluster(self):
        L = oracle_lib()
        n = self.db.n
        a32 = lambda: np.zeros(n, dtype=np.uint32)
        a64 = lambda: np.zeros(n, dtype=np.uint64)
        self.swarmid, self.generation, self.parent, self.next = a32(), a32(), a32(), a32()
        self.sw_seed, self.sw_last, self.sw_size, self.sw_singletons, self.sw_maxgen = a32(), a32(), a32(), a32(), a32()
        self.sw_mass, self.sw_sumlen = a64(), a64()
        net = self.net if self.net.size else np.zeros(1, dtype=np.uint32)
        self.nswarms = L.orc_d1_cluster(C.byref(self.c), _p(self.link_start, _u32p), _p(self.link_count, _u32p), _p(net, _u32p),
                                        _p(self.swarmid, _u32p), _p(self.generation, _u32p), _p(self.parent, _u32p),
                                        _p(self.next, _u32p), _p(self.sw_seed, _u32p), _p(self.sw_last, _u32p),
                                        _p(self.sw_size, _u32p), _p(self.sw_singletons, _u32p), _p(self.sw_maxgen, _u32p),
                                        _p(self.sw_mass, _u64p), _p(self.sw_sumlen, _u64p))
        self.swarm_of = self.sw_seed[self.swarmid]
        return self.swarm_of, self.generation, self.parent

    def fastidious(self, boundary=3, bloom_bits=16):
        L = oracle_lib()
        n = self.db.n
        self.sw_attached = np.zeros(n, dtype=np.uint8)
        self.graft_cand = np.zeros(n, dtype=np.uint32)
        self.graft_raw = np.zeros(n, dtype=np.uint32)
        st = np.zeros(4, dtype=np.uint64)
        g = L.orc_d1_fastidious(C.byref(self.c), int(boundary), int(bloom_bits), int(self.nswarms), _p(self.swarmid, _u32p),
                                _p(self.next, _u32p), _p(self.sw_seed, _u32p), _p(self.sw_last, _u32p), _p(self.sw_size, _u32p),
                                _p(self.sw_singletons, _u32p), _p(self.sw_mass, _u64p), _p(self.sw_sumlen, _u64p),
                                _p(self.sw_attached, _u8p), _p(self.graft_cand, _u32p), _p(self.graft_raw, _u32p), _p(st, _u64p))
        self.fast_stats = st
        return g

    def scoring(self, m=5, p=4, g=12, e=4):
        out = (C.c_int64 * 3)()
        oracle_lib().orc_scoring(m, p, g, e, out)
        return list(out)

    def nw_diffs(self, q, t, pen):
        """differences of the reference's optimal alignment, query amplicon q vs target amplicon t"""
        L = oracle_lib()
        s = self.db.stride
        qs = np.ascontiguousarray(self.words[q * s:(q + 1) * s])
        ts = np.ascontiguousarray(self.words[t * s:(t + 1) * s])
        return L.orc_nw_diffs(_p(ts, _u64p), int(self.len[t]), _p(qs, _u64p), int(self.len[q]), pen[0], pen[1], pen[2], None)

    def dn_cluster(self, d, no_cluster_breaking=False, pen=None):
        L = oracle_lib()
        n = self.db.n
        pen = pen or self.scoring()
        cp = (C.c_int64 * 3)(*pen)
        a32 = lambda: np.zeros(n, dtype=np.uint32)
        self.order, self.swarm_of, self.generation, self.parent, self.pdiff, self.radius = a32(), a32(), a32(), a32(), a32(), a32()
        st = np.zeros(3, dtype=np.uint64)
        self.nswarms = L.orc_dn_cluster(C.byref(self.c), int(d), int(no_cluster_breaking), cp, _p(self.order, _u32p),
                                        _p(self.swarm_of, _u32p), _p(self.generation, _u32p), _p(self.parent, _u32p),
                                        _p(self.pdiff, _u32p), _p(self.radius, _u32p), _p(st, _u64p))
        self.dn_stats = st
        return self.swarm_of, self.generation, self.parent, self.pdiff

    def derep(self):
        """d=0 (oracle_d0.c): returns (rep, mass, size, singletons) laid out like the engine's outputs — sums at the
        representatives' indices — plus self.d0_seeds (clusters in output order) and self.d0_next (chains)"""
        L = oracle_lib()
        n = self.db.n
        rep, nxt, seeds, size, singles = (np.zeros(n, dtype=np.uint32) for _ in range(5))
        mass = np.zeros(n, dtype=np.uint64)
        k = L.orc_d0_dereplicate(C.byref(self.c), _p(rep, _u32p), _p(nxt, _u32p), _p(seeds, _u32p), _p(mass, _u64p),
                                 _p(size, _u32p), _p(singles, _u32p))
        self.d0_seeds, self.d0_next, self.d0_clusters = seeds[:k].copy(), nxt, k
        m2, s2, g2 = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        m2[seeds[:k]], s2[seeds[:k]], g2[seeds[:k]] = mass[:k], size[:k], singles[:k]
        return rep, m2, s2, g2

    def swarm_lists(self):
        """final member lists (list order) of the non-attached swarms, following `next`"""
        out = []
        att = getattr(self, "sw_attached", None)
        for s in range(self.nswarms):
            if att is not None and att[s]:
                continue
            a = int(self.sw_seed[s])
            cur = []
            while a != NONE:
                cur.append(a)
                a = int(self.next[a])
            out.append(cur)
        return out


_gen = None


def gen_lib():
    global _gen
    if _gen is None:
        L = C.CDLL(str(ROOT / "tools" / "libgen_amplicons.so"))
        L.gen_create.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_double]
        L.gen_create.restype = C.c_void_p
        L.gen_count.argtypes = [C.c_void_p]
        L.gen_count.restype = C.c_uint64
        L.gen_write_fasta.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64]
        L.gen_free.argtypes = [C.c_void_p]
        L.gen_free.restype = None
        _gen = L
    return _gen


def make_fasta(path, n, L, seed=42, ab_mode=0, orphan_p=0.2):
    G = gen_lib()
    g = G.gen_create(int(n), int(L), int(seed), int(ab_mode), float(orphan_p))
    assert g
    rc = G.gen_write_fasta(g, str(path).encode(), 0, int(n))
    G.gen_free(g)
    assert rc == 0
    return path


def canonical(text: bytes) -> bytes:
    """sort ids inside each line, then sort lines"""
    lines = [b" ".join(sorted(l.split())) for l in text.splitlines() if l.strip()]
    return b"\n".join(sorted(lines)) + b"\n"


def have_ref() -> bool:
    return REF_BIN.exists() and os.access(REF_BIN, os.X_OK)


REF_CHECKS = {"ran": 0, "skipped": 0}


def with_ref() -> bool:
    """gate of the comparisons against the reference binary INSIDE a test: counts them, so that the terminal summary says how
    many ran and how many were skipped for lack of oracle/_ref/swarm (a silent pass would hide a missing binary)"""
    if have_ref():
        REF_CHECKS["ran"] += 1
        return True
    REF_CHECKS["skipped"] += 1
    return False


def make_variant_fasta(path, n, L, seed, kmin=3, kmax=60):
    """one abundant seed + n distinct variants, each kmin..kmax random edits (60 % substitutions, 20 % deletions, 20 %
    insertions) away from it: related sequences far apart, the input for large-d tests"""
    import random
    rng = random.Random(seed)
    root = "".join(rng.choice("ACGT") for _ in range(L))
    recs, seen = [("seed", 100000, root)], {root}
    while len(recs) <= n:
        s = list(root)
        for _ in range(rng.randint(kmin, kmax)):
            r, p = rng.random(), rng.randrange(len(s))
            if r < 0.6:
                s[p] = rng.choice("ACGT")
            elif r < 0.8:
                del s[p]
            else:
                s.insert(p, rng.choice("ACGT"))
        s = "".join(s)
        if s not in seen and len(s) >= 20:
            seen.add(s)
            recs.append((f"m{len(recs)}", rng.choice([1, 1, 2, 5]), s))
    Path(path).write_text("".join(f">{h}_{a}\n{t}\n" for h, a, t in recs))
    return path


def run_ref(fasta, *flags, outputs=("o",), threads=1):
    """run the unmodified reference binary; returns {flag: bytes}"""
    res = {}
    with tempfile.TemporaryDirectory() as td:
        cmd = [str(REF_BIN), "-t", str(threads), "-l", os.path.join(td, "log")]
        for o in outputs:
            cmd += ["-" + o, os.path.join(td, o)]
        if "o" not in outputs:
            cmd += ["-o", os.devnull]
        cmd += list(flags) + [str(fasta)]
        p = subprocess.run(cmd, capture_output=True)
        res["rc"] = p.returncode
        res["stderr"] = p.stderr
        for o in outputs:
            f = os.path.join(td, o)
            res[o] = open(f, "rb").read() if os.path.exists(f) else b""
        res["log"] = open(os.path.join(td, "log"), "rb").read() if os.path.exists(os.path.join(td, "log")) else b""
    return res


# ---- parity at BASELINE scale: SHA-256 of the reference's outputs (tests/golden/scale_hashes.json, written by
# tests/golden/make_scale_hashes.py in the build container, where /root/reference compiles).  A case = generator
# arguments + reference flags; bench.py and the slow GPU tests rebuild the same FASTA and compare hashes.
SCALE_CASES = {
    # name: (n, L, seed, ab_mode, reference flags, reference threads)
    "c2": (10_000_000, 150, 42, 0, (), 8),                     # BASELINE configs[1]
    "c3": (10_000_000, 150, 42, 0, ("-f",), 1),                # configs[2]; -t 1: the reference's light pass races
    "c4": (1_000_000, 400, 42, 0, ("-d", "2"), 8),             # configs[3]
    "tie1m": (1_000_000, 150, 42, 1, (), 8),                   # tie-heavy: 70 % of the abundances are 1
    "tie1m_f": (1_000_000, 150, 42, 1, ("-f",), 1),
    "c2_1m": (1_000_000, 150, 42, 0, (), 8),                   # the same generator stream at the size the GPU test suite uses
    "c3_1m": (1_000_000, 150, 42, 0, ("-f",), 1),
    "c4_100k": (100_000, 400, 42, 0, ("-d", "2"), 8),
}
SCALE_HASHES = GOLDEN / "scale_hashes.json"

_canon = None


def canon_lib():
    global _canon
    if _canon is None:
        L = C.CDLL(str(ROOT / "tools" / "libcanon.so"))
        L.canon_sha256.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        L.canon_sha256.restype = C.c_int64
        L.sha256_hex.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        L.sha256_hex.restype = None
        _canon = L
    return _canon


def canon_sha256(text: bytes) -> str:
    """SHA-256 of the canonical form (ids sorted inside each line, lines sorted)"""
    b = C.create_string_buffer(65)
    if canon_lib().canon_sha256(text, len(text), b) < 0:
        raise MemoryError("canon_sha256")
    return b.value.decode()


def sha256_hex(data: bytes) -> str:
    b = C.create_string_buffer(65)
    canon_lib().sha256_hex(data, len(data), b)
    return b.value.decode()


def scale_fasta(name: str) -> str:
    n, L, seed, ab_mode, _flags, _t = SCALE_CASES[name]
    path = f"/dev/shm/swb200_{n}x{L}_s{seed}" + ("_tie" if ab_mode else "") + ".fa"
    if not os.path.exists(path):
        make_fasta(path + ".tmp", n, L, seed, ab_mode)
        os.replace(path + ".tmp", path)
    return path


def scale_hashes() -> dict:
    import json
    return json.loads(SCALE_HASHES.read_text()) if SCALE_HASHES.exists() else {}


def output_hashes(o: bytes, s: bytes | None = None, i: bytes | None = None) -> dict:
    h = {"o_canonical_sha256": canon_sha256(o), "o_sha256": sha256_hex(o), "swarms": o.count(b"\n")}
    if s is not None:
        h["s_sha256"] = sha256_hex(s)
    if i is not None:
        h["i_sha256"] = sha256_hex(i)
    return h


def engine_case_outputs(name: str, device: int = 0, db=None, **opt):
    """the CUDA engine (through the C ABI) on SCALE_CASES[name]: returns the digests of its -o / -s / -i texts, computed
    exactly like tests/golden/make_scale_hashes.py computes the reference's"""
    from swarm_b200 import D1Result, DnResult, Engine, HostDb, scoring
    n, L, seed, ab_mode, flags, _t = SCALE_CASES[name]
    d = int(flags[flags.index("-d") + 1]) if "-d" in flags else 1
    own_db = db is None
    if own_db:
        db = HostDb(scale_fasta(name), check_dup_sequences=d > 1)
    eng = Engine(device, **opt)
    try:
        eng.load(db)
        if d == 1:
            eng.d1_index()
            eng.d1_network()
            sw, gen, par = eng.d1_cluster()
            gc = None
            if "-f" in flags:
                gc, _nl, _nh = eng.d1_fastidious(boundary=3)
            res = D1Result(db, sw, gen, par, graft_cand=gc, boundary=3)
        else:
            sw, gen, par, pd = eng.dn_cluster(d, penalties=scoring())
            res = DnResult(db, sw, gen, par, pd)
        h = output_hashes(res.swarms_text(), res.stats_text(), res.structure_text())
        res.close()
    finally:
        eng.close()
        if own_db:
            db.close()
    return h


def compare_case(name: str, got: dict) -> list:
    """keys on which `got` differs from the committed reference digests of case `name` ([] = parity)"""
    want = scale_hashes().get(name)
    if want is None:
        return ["no golden digests for " + name]
    return [k for k in ("o_canonical_sha256", "o_sha256", "s_sha256", "i_sha256", "swarms") if got.get(k) != want.get(k)]
