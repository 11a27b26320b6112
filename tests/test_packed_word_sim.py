"""CPU check of the argument behind the packed relaxation word of the clustering kernels (csrc/d1_kernels.cuh: k_cluster_persistent<.., PACK>,
csrc/d1_bucket.cuh): relaxing ONE word `swarm | generation | parent` with offers `((word[u] | idmask) + 1) | u` under an ARBITRARY
relaxation order reaches the same fixed point as the reference semantics — key[v] = min over links u->v of key[u] + 1 with
key = swarm << 32 | generation, then parent[v] = min {u : key[u] + 1 == key[v]} (closed form of src/algod1.cc:1185-1280, :673-718) — and a
generation that does not fit the word is always detected at unpack time (a final generation field of all ones), never silently wrong."""
import random


def unpacked(n, edges):
    key = [v << 32 for v in range(n)]
    changed = True
    while changed:
        changed = False
        for u, v in edges:
            c = key[u] + 1
            if c < key[v]:
                key[v] = c
                changed = True
    par = [None] * n
    for u, v in edges:
        if key[u] + 1 == key[v] and (par[v] is None or u < par[v]):
            par[v] = u
    return [k >> 32 for k in key], [k & 0xFFFFFFFF for k in key], par


def packed(n, edges, ib, gb, rng, stale=False):
    """stale=True models the fire-and-forget kernels: the offer is computed from, and "lowered" is decided against, words that were
    read EARLIER (here: at any point since the start of the round — another thread may have lowered them since), and the atomicMin
    returns nothing"""
    idm, gm = (1 << ib) - 1, (1 << gb) - 1
    w = [(v << (gb + ib)) | idm for v in range(n)]
    lowered = set(range(n))
    while lowered:                                   # rounds: only links whose source KEY went down last round are offered again
        nxt = set()
        es = [e for e in edges if e[0] in lowered]
        rng.shuffle(es)                              # any order inside a round (the GPU's is arbitrary)
        start = list(w)
        for u, v in es:
            src = start[u] if stale and rng.random() < 0.5 else w[u]      # a source word read before a concurrent lowering
            seen = start[v] if stale and rng.random() < 0.5 else w[v]     # the destination word loaded before the atomic
            cand = ((src | idm) + 1) | u
            if cand < seen:
                if stale:
                    w[v] = min(w[v], cand)           # RED.MIN: no return value
                    if (seen | idm) > (cand | idm):
                        nxt.add(v)                   # superset: v may have been lowered by somebody else this round
                else:
                    old, w[v] = w[v], cand
                    if (old | idm) > (cand | idm):
                        nxt.add(v)
        # (a source whose stale word was offered was lowered in this round, so the offer that lowered it marked it: it re-offers)
        lowered = nxt
    deep = any(((x >> ib) & gm) == gm for x in w)
    return [x >> (gb + ib) for x in w], [(x >> ib) & gm for x in w], [None if (x & idm) == idm else x & idm for x in w], deep


def test_packed_word_equals_key_plus_parent_pass_and_detects_overflow():
    rng = random.Random(1)
    seen_deep = seen_exact = 0
    for _trial in range(300):
        n = rng.randint(2, 60)
        edges = []
        for _ in range(rng.randint(0, 3 * n)):
            a, b = rng.randrange(n), rng.randrange(n)
            if a == b:
                continue
            edges += [(a, b), (b, a)] if rng.random() < 0.3 else [(min(a, b), max(a, b))]      # ties link both ways
        if rng.random() < 0.3:
            edges += [(i, i + 1) for i in range(n - 1)]                                          # a deep chain
        ib, gb = n.bit_length(), rng.choice([2, 3, 4, 16])
        s, g, p = unpacked(n, edges)
        for stale in (False, True):
            s2, g2, p2, deep = packed(n, edges, ib, gb, rng, stale)
            if max(g) >= (1 << gb) - 1:
                assert deep, "a generation beyond the field went undetected"
                seen_deep += 1
            else:
                assert not deep and (s, g, p) == (s2, g2, p2), stale
                seen_exact += 1
    assert seen_deep > 40 and seen_exact > 200
