"""ctypes bindings of include/swarm_b200.h and include/swarm_b200_host.h (no algorithm code here)."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
ENUM_FULL, ENUM_HALF, ENUM_JOIN = 0, 1, 2
NONE = 0xFFFFFFFF
_STATUS = {1: "CUDA error", 2: "invalid argument", 3: "duplicate sequences", 4: "out of memory", 5: "unsupported"}


class EngineError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"swarm_b200 status {status} ({_STATUS.get(status, '?')}): {text}")
        self.status = status
        self.text = text


def lib_paths() -> dict:
    return {"engine": _HERE / "libswarm_b200.so", "host": _HERE / "libswarm_b200_host.so"}


_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)


def _ptr(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


_engine_lib = None
_host_lib = None


def engine_lib():
    """Load libswarm_b200.so (fails loudly if it has not been built)."""
    global _engine_lib
    if _engine_lib is None:
        p = lib_paths()["engine"]
        if not p.exists():
            raise FileNotFoundError(f"{p} is missing: run `make engine` (or __graft_entry__.build())")
        L = C.CDLL(str(p))
        vp = C.c_void_p
        L.swb200_last_error.restype = C.c_char_p
        L.swb200_create.argtypes = [C.POINTER(vp), C.c_int]
        L.swb200_destroy.argtypes = [vp]
        L.swb200_destroy.restype = None
        L.swb200_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        L.swb200_load_db.argtypes = [vp, _u64p, C.c_uint32, _u32p, _u64p, C.c_uint32]
        L.swb200_load_db_compact.argtypes = [vp, _u64p, C.c_uint32, C.POINTER(C.c_uint16), _u64p, _u32p, C.c_uint32, C.c_uint32]
        L.swb200_load_db_shard.argtypes = [vp, _u64p, C.c_uint32, _u32p, _u64p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.swb200_load_db_shard_compact.argtypes = [vp, _u64p, C.c_uint32, C.POINTER(C.c_uint16), C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _u32p, C.c_uint32]
        L.swb200_load_db_rows.argtypes = [vp, _u64p, C.c_uint32, _u32p, _u64p, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, C.c_uint32]
        L.swb200_db_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.swb200_db_commit.argtypes = [vp]
        L.swb200_load_db_device.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_uint32]
        L.swb200_d1_get_cluster.argtypes = [vp, C.c_uint32, C.c_uint32, _u32p, _u32p, _u32p]
        L.swb200_dist_buffer_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.swb200_dist_buffer_bytes.restype = C.c_uint64
        L.swb200_dist_row_count.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.swb200_dist_row_count.restype = C.c_uint32
        L.swb200_dist_row_id.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.swb200_dist_row_id.restype = C.c_uint32
        L.swb200_dist_setup.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(vp), C.c_uint64]
        L.swb200_d1_cluster_dist.argtypes = [vp, _u32p, _u32p, _u32p]
        L.swb200_d1_reserve.argtypes = [vp]
        L.swb200_d1_index.argtypes = [vp]
        L.swb200_d1_network.argtypes = [vp, C.c_int, _u64p]
        L.swb200_d1_get_network.argtypes = [vp, _u64p, _u32p]
        L.swb200_d1_export_links.argtypes = [vp, _u32p]
        L.swb200_d1_import_links.argtypes = [vp, _u32p, C.c_uint64]
        L.swb200_d1_links_device.argtypes = [vp, C.POINTER(vp), _u64p]
        L.swb200_d1_import_links_device.argtypes = [vp, vp, C.c_uint64]
        L.swb200_d1_cluster.argtypes = [vp, _u32p, _u32p, _u32p]
        L.swb200_d1_fastidious.argtypes = [vp, C.c_uint64, _u32p, _u64p, _u64p]
        L.swb200_dn_cluster.argtypes = [vp, C.c_uint32, C.c_int, C.POINTER(C.c_int64), _u32p, _u32p, _u32p, _u32p]
        L.swb200_d0_dereplicate.argtypes = [vp, _u32p, _u64p, _u32p, _u32p, _u64p]
        L.swb200_stream.argtypes = [vp, C.POINTER(vp)]
        L.swb200_last_device_seconds.argtypes = [vp]
        L.swb200_last_device_seconds.restype = C.c_double
        L.swb200_phase_device_seconds.argtypes = [vp, C.c_int]
        L.swb200_phase_device_seconds.restype = C.c_double
        L.swb200_get_stats.argtypes = [vp, _u64p, C.c_int]
        L.swb200_debug_variants.argtypes = [vp, C.c_uint32, C.c_int, _u64p, _u32p, C.c_uint32, _u32p, _u64p, C.c_uint32, _u32p]
        _engine_lib = L
    return _engine_lib


def host_lib():
    global _host_lib
    if _host_lib is None:
        p = lib_paths()["host"]
        if not p.exists():
            raise FileNotFoundError(f"{p} is missing: run `make host` (or __graft_entry__.build())")
        L = C.CDLL(str(p))
        vp = C.c_void_p
        L.swbh_last_error.restype = C.c_char_p
        L.swbh_db_read_fasta.argtypes = [C.c_char_p, C.c_int, C.c_int64, C.c_int, C.POINTER(vp)]
        L.swbh_db_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int64, C.c_int, C.POINTER(vp)]
        L.swbh_set_threads.argtypes = [C.c_int]
        L.swbh_set_threads.restype = None
        L.swbh_set_writer_grain.argtypes = [C.c_uint64]
        L.swbh_set_writer_grain.restype = None
        L.swbh_db_free.argtypes = [vp]
        L.swbh_db_free.restype = None
        for f in ("count", "longest", "stride_words"):
            getattr(L, f"swbh_db_{f}").argtypes = [vp]
            getattr(L, f"swbh_db_{f}").restype = C.c_uint32
        L.swbh_db_nucleotides.argtypes = [vp]
        L.swbh_db_nucleotides.restype = C.c_uint64
        L.swbh_db_words.argtypes = [vp]
        L.swbh_db_words.restype = _u64p
        L.swbh_db_lengths.argtypes = [vp]
        L.swbh_db_lengths.restype = _u32p
        L.swbh_db_abundances.argtypes = [vp]
        L.swbh_db_abundances.restype = _u64p
        L.swbh_db_header.argtypes = [vp, C.c_uint32]
        L.swbh_db_header.restype = C.c_char_p
        L.swbh_d1_assemble.argtypes = [vp, _u32p, _u32p, _u32p, _u32p, C.c_uint64, C.POINTER(vp)]
        L.swbh_result_free.argtypes = [vp]
        L.swbh_result_free.restype = None
        L.swbh_result_swarms.argtypes = [vp]
        L.swbh_result_swarms.restype = C.c_uint64
        L.swbh_result_grafts.argtypes = [vp]
        L.swbh_result_grafts.restype = C.c_uint64
        L.swbh_result_largest.argtypes = [vp]
        L.swbh_result_largest.restype = C.c_uint32
        L.swbh_result_maxgen.argtypes = [vp]
        L.swbh_result_maxgen.restype = C.c_uint32
        cpp = C.POINTER(C.c_char_p)
        L.swbh_write_swarms.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int, C.c_int64, C.POINTER(vp), _u64p]
        L.swbh_write_stats.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_write_structure.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_write_seeds.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_write_uclust.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int64), C.c_int, C.c_int64, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_uclust_band.argtypes = [C.c_int]
        L.swbh_uclust_band.restype = None
        L.swbh_write_network.argtypes = [vp, _u64p, _u32p, C.c_int, C.c_int64, C.POINTER(vp), _u64p]
        L.swbh_dn_assemble.argtypes = [vp, _u32p, _u32p, _u32p, _u32p, C.POINTER(vp)]
        L.swbh_dn_write_stats.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_dn_write_structure.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_d0_assemble.argtypes = [vp, _u32p, _u64p, _u32p, _u32p, C.POINTER(vp)]
        L.swbh_derep_free.argtypes = [vp]
        L.swbh_derep_free.restype = None
        L.swbh_derep_clusters.argtypes = [vp]
        L.swbh_derep_clusters.restype = C.c_uint64
        L.swbh_derep_largest.argtypes = [vp]
        L.swbh_derep_largest.restype = C.c_uint32
        L.swbh_derep_heaviest.argtypes = [vp]
        L.swbh_derep_heaviest.restype = C.c_uint64
        L.swbh_d0_write_swarms.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int64, C.POINTER(vp), _u64p]
        L.swbh_d0_write_seeds.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_d0_write_uclust.argtypes = [vp, vp, C.c_int, C.c_int64, C.POINTER(vp), _u64p]
        L.swbh_d0_write_structure.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_d0_write_stats.argtypes = [vp, vp, C.c_int, C.POINTER(vp), _u64p]
        L.swbh_scoring.argtypes = [C.c_int64] * 4 + [C.POINTER(C.c_int64)]
        L.swbh_scoring.restype = None
        L.swbh_free.argtypes = [vp]
        L.swbh_free.restype = None
        del cpp
        _host_lib = L
    return _host_lib


def compact_form(lengths, abundance):
    """(len16, run_abundance, run_start) of a database sorted by abundance: the arguments of swb200_load_db_compact"""
    lengths = np.asarray(lengths)
    abundance = np.ascontiguousarray(abundance, dtype=np.uint64)
    assert lengths.max(initial=0) <= 65535
    starts = np.flatnonzero(np.concatenate(([True], abundance[1:] != abundance[:-1]))).astype(np.uint32)
    return lengths.astype(np.uint16), abundance[starts].copy(), np.concatenate((starts, [abundance.shape[0]])).astype(np.uint32)


class HostDb:
    """The sorted, 2-bit packed amplicon database (host mirror of the reference's db.cc)."""

    def __init__(self, path: str | os.PathLike | None = None, text: bytes | None = None, usearch_abundance=False,
                 append_abundance=0, check_dup_sequences=False):
        L = host_lib()
        h = C.c_void_p()
        self.opts = (int(bool(usearch_abundance)), int(append_abundance))
        if text is not None:
            rc = L.swbh_db_parse(text, len(text), self.opts[0], self.opts[1], int(check_dup_sequences), C.byref(h))
        else:
            rc = L.swbh_db_read_fasta(str(path).encode(), self.opts[0], self.opts[1], int(check_dup_sequences), C.byref(h))
        if rc != 0:
            raise ValueError(L.swbh_last_error().decode())
        self._h = h
        self.n = L.swbh_db_count(h)
        self.longest = L.swbh_db_longest(h)
        self.stride = L.swbh_db_stride_words(h)
        self.nucleotides = L.swbh_db_nucleotides(h)
        self.words = np.ctypeslib.as_array(L.swbh_db_words(h), shape=(self.n * self.stride,))
        self.len = np.ctypeslib.as_array(L.swbh_db_lengths(h), shape=(self.n,))
        self.abundance = np.ctypeslib.as_array(L.swbh_db_abundances(h), shape=(self.n,))

    def header(self, i: int) -> str:
        return host_lib().swbh_db_header(self._h, int(i)).decode()

    def headers(self):
        L = host_lib()
        return [L.swbh_db_header(self._h, i).decode() for i in range(self.n)]

    def close(self):
        if getattr(self, "_h", None):
            self.words = self.len = self.abundance = None
            host_lib().swbh_db_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class D1Result:
    """Swarm lists / stats / structure built by the host writers from the engine's arrays."""

    def __init__(self, db: HostDb, swarm_of, generation, parent, graft_cand=None, boundary=3):
        L = host_lib()
        self.db = db
        self._keep = [np.ascontiguousarray(a, dtype=np.uint32) for a in (swarm_of, generation, parent)]
        gc = np.ascontiguousarray(graft_cand, dtype=np.uint32) if graft_cand is not None else None
        h = C.c_void_p()
        rc = L.swbh_d1_assemble(db._h, *[_ptr(a, _u32p) for a in self._keep], _ptr(gc, _u32p), int(boundary), C.byref(h))
        if rc != 0:
            raise ValueError(L.swbh_last_error().decode())
        self._h = h
        self.swarms = L.swbh_result_swarms(h)
        self.grafts = L.swbh_result_grafts(h)
        self.largest = L.swbh_result_largest(h)
        self.maxgen = L.swbh_result_maxgen(h)

    def _text(self, fn, *args) -> bytes:
        L = host_lib()
        out = C.c_void_p()
        n = C.c_uint64()
        rc = fn(*args, C.byref(out), C.byref(n))
        if rc != 0:
            raise ValueError(L.swbh_last_error().decode())
        data = C.string_at(out, n.value)
        L.swbh_free(out)
        return data

    def swarms_text(self, mothur=False, differences=1) -> bytes:
        L = host_lib()
        return self._text(L.swbh_write_swarms, self.db._h, self._h, int(mothur), int(differences), *self.db.opts)

    def stats_text(self) -> bytes:
        L = host_lib()
        return self._text(L.swbh_write_stats, self.db._h, self._h, self.db.opts[0])

    def structure_text(self) -> bytes:
        L = host_lib()
        return self._text(L.swbh_write_structure, self.db._h, self._h, self.db.opts[0])

    def seeds_text(self) -> bytes:
        L = host_lib()
        return self._text(L.swbh_write_seeds, self.db._h, self._h, self.db.opts[0])

    def uclust_text(self, differences=1, penalties=(18, 24, 13), threads=1) -> bytes:
        """`-u` records; penalties = converted costs (mismatch, gap open, gap extension), defaults = swarm's defaults"""
        L = host_lib()
        pen = (C.c_int64 * 3)(*[int(x) for x in penalties])
        return self._text(L.swbh_write_uclust, self.db._h, self._h, int(differences), pen, *self.db.opts, int(threads))

    def close(self):
        if getattr(self, "_h", None):
            host_lib().swbh_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DerepResult:
    """d=0: clusters of identical sequences in the reference's output order, built by the host layer from the
    engine's rep / mass / size / singletons arrays (swb200_d0_dereplicate)."""

    def __init__(self, db: HostDb, rep, mass, size, singletons):
        L = host_lib()
        self.db = db
        rep = np.ascontiguousarray(rep, dtype=np.uint32)
        mass = np.ascontiguousarray(mass, dtype=np.uint64)
        size = np.ascontiguousarray(size, dtype=np.uint32)
        singletons = np.ascontiguousarray(singletons, dtype=np.uint32)
        h = C.c_void_p()
        if L.swbh_d0_assemble(db._h, _ptr(rep, _u32p), _ptr(mass, _u64p), _ptr(size, _u32p), _ptr(singletons, _u32p), C.byref(h)) != 0:
            raise ValueError(L.swbh_last_error().decode())
        self._h = h
        self.clusters = L.swbh_derep_clusters(h)
        self.largest = L.swbh_derep_largest(h)
        self.heaviest = L.swbh_derep_heaviest(h)

    _text = D1Result._text

    def swarms_text(self, mothur=False) -> bytes:
        return self._text(host_lib().swbh_d0_write_swarms, self.db._h, self._h, int(mothur), *self.db.opts)

    def seeds_text(self) -> bytes:
        return self._text(host_lib().swbh_d0_write_seeds, self.db._h, self._h, self.db.opts[0])

    def uclust_text(self) -> bytes:
        return self._text(host_lib().swbh_d0_write_uclust, self.db._h, self._h, *self.db.opts)

    def structure_text(self) -> bytes:
        return self._text(host_lib().swbh_d0_write_structure, self.db._h, self._h, self.db.opts[0])

    def stats_text(self) -> bytes:
        return self._text(host_lib().swbh_d0_write_stats, self.db._h, self._h, self.db.opts[0])

    def close(self):
        if getattr(self, "_h", None):
            host_lib().swbh_derep_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DnResult(D1Result):
    """d>1 flavour: same swarm lists; stats carry max radius, structure lines carry the differences."""

    def __init__(self, db: HostDb, swarm_of, generation, parent, pdiff):
        L = host_lib()
        self.db = db
        self._keep = [np.ascontiguousarray(a, dtype=np.uint32) for a in (swarm_of, generation, parent, pdiff)]
        h = C.c_void_p()
        rc = L.swbh_dn_assemble(db._h, *[_ptr(a, _u32p) for a in self._keep], C.byref(h))
        if rc != 0:
            raise ValueError(L.swbh_last_error().decode())
        self._h = h
        self.swarms = L.swbh_result_swarms(h)
        self.grafts = 0
        self.largest = L.swbh_result_largest(h)
        self.maxgen = max(1, L.swbh_result_maxgen(h))

    def stats_text(self) -> bytes:
        return self._text(host_lib().swbh_dn_write_stats, self.db._h, self._h, self.db.opts[0])

    def structure_text(self) -> bytes:
        return self._text(host_lib().swbh_dn_write_structure, self.db._h, self._h, self.db.opts[0])


def dist_buffer_bytes(n_total: int, world: int, items_per_amplicon: int = 4) -> int:
    return int(engine_lib().swb200_dist_buffer_bytes(int(n_total), int(world), int(items_per_amplicon)))


def dist_row_ids(n_total: int, rank: int, world: int) -> np.ndarray:
    """amplicon ids of the rows swb200_d1_cluster_dist returns on `rank` (block-cyclic, blocks of 4096 ids)"""
    L = engine_lib()
    rows = int(L.swb200_dist_row_count(int(n_total), int(rank), int(world)))
    blk = 4096
    assert int(L.swb200_dist_row_id(1, 3, blk + 5)) == (1 * 3 + 1) * blk + 5      # the engine's block size
    i = np.arange(rows, dtype=np.int64)
    return (((i // blk) * world + rank) * blk + (i % blk)).astype(np.uint32)


def scoring(match=5, mismatch=4, gap_open=12, gap_extend=4):
    out = (C.c_int64 * 3)()
    host_lib().swbh_scoring(match, mismatch, gap_open, gap_extend, out)
    return list(out)


def network_text(db: HostDb, row_ptr, col) -> bytes:
    L = host_lib()
    out = C.c_void_p()
    n = C.c_uint64()
    rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    cc = np.ascontiguousarray(col, dtype=np.uint32)
    rc = L.swbh_write_network(db._h, _ptr(rp, _u64p), _ptr(cc, _u32p), *db.opts, C.byref(out), C.byref(n))
    if rc != 0:
        raise ValueError(L.swbh_last_error().decode())
    data = C.string_at(out, n.value)
    L.swbh_free(out)
    return data


class Engine:
    """One CUDA context of the engine (one per GPU / per rank)."""

    def __init__(self, device: int = 0, **options):
        L = engine_lib()
        h = C.c_void_p()
        rc = L.swb200_create(C.byref(h), int(device))
        if rc != 0:
            raise EngineError(rc, L.swb200_last_error().decode())
        self._h = h
        self.n = 0
        for k, v in options.items():
            self.set_option(k, v)

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(rc, engine_lib().swb200_last_error().decode())

    def set_option(self, key: str, value: int):
        self._ck(engine_lib().swb200_set_option(self._h, key.encode(), int(value)))

    def load_db(self, words, stride, lengths, abundance):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        abundance = np.ascontiguousarray(abundance, dtype=np.uint64)
        n = lengths.shape[0]
        assert words.shape[0] == n * stride and abundance.shape[0] == n
        self._ck(engine_lib().swb200_load_db(self._h, _ptr(words, _u64p), int(stride), _ptr(lengths, _u32p),
                                            _ptr(abundance, _u64p), n))
        self.n = n

    def load_db_compact(self, words, stride, len16, run_abundance, run_start):
        """swb200_load_db_compact: u16 lengths + abundance runs (see compact_form)"""
        n = len16.shape[0]
        assert words.shape[0] == n * stride and len16.dtype == np.uint16 and run_start.shape[0] == run_abundance.shape[0] + 1
        self._ck(engine_lib().swb200_load_db_compact(self._h, _ptr(words, _u64p), int(stride), len16.ctypes.data_as(C.POINTER(C.c_uint16)),
                                                    _ptr(run_abundance, _u64p), _ptr(run_start, _u32p), int(run_abundance.shape[0]), n))
        self.n = n

    def load_db_shard(self, words, stride, lengths, abundance, n_total, first):
        """this rank's rows [first, first+len(lengths)) of a database of n_total amplicons; then exchange + db_commit"""
        count = lengths.shape[0]
        assert words.shape[0] == count * stride and abundance.shape[0] == count
        self._ck(engine_lib().swb200_load_db_shard(self._h, _ptr(words, _u64p), int(stride), _ptr(lengths, _u32p),
                                                  _ptr(abundance, _u64p), int(n_total), int(first), int(count)))
        self.n = int(n_total)

    def load_db_shard_compact(self, words, stride, len16, n_total, first, run_abundance, run_start):
        """load_db_shard from fewer host bytes: u16 lengths of this rank's rows, abundance runs of the WHOLE database (no abundance array)"""
        count = len16.shape[0]
        assert words.shape[0] == count * stride and len16.dtype == np.uint16 and run_start.shape[0] == run_abundance.shape[0] + 1
        self._ck(engine_lib().swb200_load_db_shard_compact(self._h, _ptr(words, _u64p), int(stride), _ptr(len16, C.POINTER(C.c_uint16)), int(n_total),
                                                          int(first), int(count), _ptr(run_abundance, _u64p), _ptr(run_start, _u32p),
                                                          int(run_abundance.shape[0])))
        self.n = int(n_total)

    def load_db_rows(self, words, stride, lengths, abundance, n_total, first, run_start):
        """sharded database: this context keeps only the rows [first, first+len(lengths)); run_start = abundance runs of the whole job"""
        count = lengths.shape[0]
        run_start = np.ascontiguousarray(run_start, dtype=np.uint32)
        assert words.shape[0] == count * stride and abundance.shape[0] == count
        self._ck(engine_lib().swb200_load_db_rows(self._h, _ptr(words, _u64p), int(stride), _ptr(lengths, _u32p), _ptr(abundance, _u64p),
                                                 int(n_total), int(first), int(count), _ptr(run_start, _u32p), int(run_start.shape[0] - 1)))
        self.n = int(n_total)

    def db_device(self):
        w, l, a = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(engine_lib().swb200_db_device(self._h, C.byref(w), C.byref(l), C.byref(a)))
        return w.value, l.value, a.value

    def db_commit(self):
        self._ck(engine_lib().swb200_db_commit(self._h))

    def load_db_device(self, d_words: int, stride: int, d_len: int, d_abundance: int, n: int):
        self._ck(engine_lib().swb200_load_db_device(self._h, C.c_void_p(d_words), int(stride), C.c_void_p(d_len),
                                                   C.c_void_p(d_abundance), int(n)))
        self.n = int(n)

    def d1_get_cluster(self, first, count, out):
        self._ck(engine_lib().swb200_d1_get_cluster(self._h, int(first), int(count), _ptr(out["swarm_of"], _u32p),
                                                   _ptr(out["generation"], _u32p), _ptr(out["parent"], _u32p)))
        return out["swarm_of"], out["generation"], out["parent"]

    def dist_setup(self, rank: int, world: int, peer_ptrs, nbytes: int):
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        self._ck(engine_lib().swb200_dist_setup(self._h, int(rank), int(world), arr, int(nbytes)))

    def d1_reserve(self):
        self._ck(engine_lib().swb200_d1_reserve(self._h))

    def d1_cluster_dist(self, out=None):
        """multi-GPU clustering (every rank calls it); out: dict of caller-owned uint32 arrays for this rank's rows"""
        o = out or {}
        self._ck(engine_lib().swb200_d1_cluster_dist(self._h, _ptr(o.get("swarm_of"), _u32p), _ptr(o.get("generation"), _u32p),
                                                    _ptr(o.get("parent"), _u32p)))
        return o.get("swarm_of"), o.get("generation"), o.get("parent")

    def load(self, db: HostDb):
        self.load_db(db.words, db.stride, db.len, db.abundance)

    def d1_index(self):
        self._ck(engine_lib().swb200_d1_index(self._h))

    def d1_network(self, no_cluster_breaking=False) -> int:
        m = C.c_uint64()
        self._ck(engine_lib().swb200_d1_network(self._h, int(bool(no_cluster_breaking)), C.byref(m)))
        self.n_links = m.value
        return m.value

    def d1_get_network(self):
        rp = np.zeros(self.n + 1, dtype=np.uint64)
        col = np.zeros(max(self.n_links, 1), dtype=np.uint32)
        self._ck(engine_lib().swb200_d1_get_network(self._h, _ptr(rp, _u64p), _ptr(col, _u32p)))
        return rp, col[: self.n_links]

    def d1_export_links(self):
        pairs = np.zeros((max(self.n_links, 1), 2), dtype=np.uint32)
        self._ck(engine_lib().swb200_d1_export_links(self._h, _ptr(pairs, _u32p)))
        return pairs[: self.n_links]

    def d1_import_links(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        self._ck(engine_lib().swb200_d1_import_links(self._h, _ptr(pairs, _u32p), pairs.shape[0]))
        self.n_links = pairs.shape[0]

    def d1_links_device(self):
        p = C.c_void_p()
        m = C.c_uint64()
        self._ck(engine_lib().swb200_d1_links_device(self._h, C.byref(p), C.byref(m)))
        return p.value, m.value

    def d1_import_links_device(self, dptr: int, n_links: int):
        self._ck(engine_lib().swb200_d1_import_links_device(self._h, C.c_void_p(dptr), int(n_links)))
        self.n_links = int(n_links)

    def d1_cluster(self, want=("swarm_of", "generation", "parent"), out=None):
        """out: optional dict of caller-owned (e.g. pinned) uint32 arrays to receive the results"""
        outs = {k: ((out[k] if out and k in out else np.empty(self.n, dtype=np.uint32)) if k in want else None)
                for k in ("swarm_of", "generation", "parent")}
        self._ck(engine_lib().swb200_d1_cluster(self._h, _ptr(outs["swarm_of"], _u32p), _ptr(outs["generation"], _u32p),
                                               _ptr(outs["parent"], _u32p)))
        return outs["swarm_of"], outs["generation"], outs["parent"]

    def d1_fastidious(self, boundary=3, want=True, out=None):
        gc = (out if out is not None else np.empty(self.n, dtype=np.uint32)) if want else None
        nl, nh = C.c_uint64(), C.c_uint64()
        self._ck(engine_lib().swb200_d1_fastidious(self._h, int(boundary), _ptr(gc, _u32p), C.byref(nl), C.byref(nh)))
        return gc, nl.value, nh.value

    def dn_cluster(self, d, no_cluster_breaking=False, penalties=(18, 24, 13), want=True, out=None):
        """out: optional dict of caller-owned uint32 arrays (swarm_of, generation, parent, pdiff); want=False: no download"""
        keys = ("swarm_of", "generation", "parent", "pdiff")
        outs = [((out[k] if out and k in out else np.empty(self.n, dtype=np.uint32)) if want else None) for k in keys]
        pen = (C.c_int64 * 3)(*penalties)
        self._ck(engine_lib().swb200_dn_cluster(self._h, int(d), int(bool(no_cluster_breaking)), pen,
                                               *[_ptr(a, _u32p) for a in outs]))
        return tuple(outs)

    def d0_dereplicate(self):
        """d=0: (rep, mass, size, singletons, n_clusters); the sums live at the representatives' indices"""
        rep, size, singles = (np.empty(self.n, dtype=np.uint32) for _ in range(3))
        mass = np.empty(self.n, dtype=np.uint64)
        k = C.c_uint64()
        self._ck(engine_lib().swb200_d0_dereplicate(self._h, _ptr(rep, _u32p), _ptr(mass, _u64p), _ptr(size, _u32p),
                                                   _ptr(singles, _u32p), C.byref(k)))
        return rep, mass, size, singles, k.value

    def stream(self) -> int:
        """cudaStream_t of the engine (wrap with torch.cuda.ExternalStream to record events / order collectives)"""
        p = C.c_void_p()
        self._ck(engine_lib().swb200_stream(self._h, C.byref(p)))
        return p.value or 0

    def last_device_seconds(self) -> float:
        return engine_lib().swb200_last_device_seconds(self._h)

    def phase_seconds(self, phase: int) -> float:
        return engine_lib().swb200_phase_device_seconds(self._h, int(phase))

    def stats(self):
        out = np.zeros(19, dtype=np.uint64)
        self._ck(engine_lib().swb200_get_stats(self._h, _ptr(out, _u64p), 19))
        return {"variants": int(out[0]), "filter_pass": int(out[1]), "slots_visited": int(out[2]),
                "exact_compares": int(out[3]), "links": int(out[4]), "launches": int(out[5]), "rows_gathered": int(out[6]), "cluster_rounds": int(out[7]),
                "fast_light_variants": int(out[8]), "fast_heavy_variants": int(out[9]),
                "fast_tag_matches": int(out[10]), "fast_verified": int(out[11]),
                "dn_qgram_comparisons": int(out[12]), "dn_alignments": int(out[13]), "dn_pruned": int(out[14]),
                "dn_links": int(out[15]), "tile_overflow": int(out[16]), "skew_fallbacks": int(out[17]),
                "cluster_unpacked_reruns": int(out[18])}

    def debug_variants(self, seed: int, mode: int, cap: int = 1 << 16):
        h = np.zeros(cap, dtype=np.uint64)
        c = np.zeros(cap, dtype=np.uint32)
        cnt = C.c_uint32()
        zl = C.c_uint32()
        z = np.zeros(4 * 8192, dtype=np.uint64)
        self._ck(engine_lib().swb200_debug_variants(self._h, int(seed), int(mode), _ptr(h, _u64p), _ptr(c, _u32p), cap,
                                                   C.byref(cnt), _ptr(z, _u64p), z.shape[0], C.byref(zl)))
        m = min(cnt.value, cap)
        return h[:m], c[:m], z[: zl.value * 4].reshape(-1, 4)

    def close(self):
        if getattr(self, "_h", None):
            engine_lib().swb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
