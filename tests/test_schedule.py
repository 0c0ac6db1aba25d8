"""Invariants of the phase table every rank's kernel walks (csrc/schedule.cc), checked without a
GPU through cdprobe_schedule().  These are the properties the device barrier and the parity of the
matrices rely on:

  * every rank has the same number of phases; the flag exchange that closes a phase is symmetric
    (i waits for j <=> j signals i), spans every rank at the end of a run, and always contains every
    rank whose transfers share an NVLink port with this rank's on either side of the barrier;
  * in a phase a rank touches at most one peer over NVLink and is touched by at most one
    (exclusive endpoints), and the pairing follows the tournament plan;
  * every ordered pair is read exactly once and written exactly once per run;
  * every landing slot that is written is verified exactly once, by its owner, strictly later,
    naming the right writer; overlapped verifies run on CTAs disjoint from the link job;
  * the table fits CDPROBE_MAX_PHASES or the call refuses.
"""
import ctypes as C
import itertools

import pytest

NONE, READ, WRITE, VERIFY, WARM = 0, 1, 2, 3, 4
UNI, SERIAL, DIAG, ALLRANK, PAIRBAR = 0x80, 0x100, 0x04, 0x400, 0x800


def schedule(pkg, n, rank, mode=1, ops=3, flags=0, ctas=148, vctas=32, nbytes=1 << 30):
    s = pkg.abi.ScheduleT()
    rc = pkg.abi.load_library().cdprobe_schedule(n, rank, nbytes, mode, ops, flags, ctas, vctas, C.byref(s))
    return rc, s


def table(pkg, n, **kw):
    out = []
    for r in range(n):
        rc, s = schedule(pkg, n, r, **kw)
        if rc != 0:
            return rc, None
        out.append(s)
    return 0, out


def slot_of(i, j):
    return i if i < j else i - 1


CASES = [dict(flags=f, ops=o, ctas=c) for f in (0, UNI, SERIAL, UNI | SERIAL, DIAG, UNI | DIAG, ALLRANK, ALLRANK | UNI, PAIRBAR)
         for o in (1, 2, 3) for c in (148, 8, 1)]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 16])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"f{c['flags']:x}-o{c['ops']}-c{c['ctas']}")
def test_schedule_invariants(pkg, n, case):
    rc, tabs = table(pkg, n, **case)
    if rc != 0:
        # only allowed when the table cannot fit: unidirectional + serial verify at large N
        assert rc == pkg.abi.ERR_ARG and (case["flags"] & UNI) and n >= 13
        return
    flags, ops, ctas = case["flags"], case["ops"], case["ctas"]
    plan = pkg.plan(n, 1 << 30, 1, flags & DIAG)
    np_ = tabs[0].n_phases
    assert np_ <= pkg.abi.MAX_PHASES
    # same length on every rank; the last barrier spans all ranks; the exchange is symmetric
    everyone = [((1 << n) - 1) & ~(1 << r) for r in range(n)]
    for r, t in enumerate(tabs):
        assert t.n_phases == np_ and t.peer_mask == everyone[r]
        for ph in range(np_):
            assert t.sync_mask[ph] & ~everyone[r] == 0 and t.post_mask[ph] & ~everyone[r] == 0
            assert t.sync_mask[ph] & t.post_mask[ph] == 0
            assert t.sync_all[ph] == (1 if n > 1 and t.sync_mask[ph] == everyone[r] else 0)
            for j in range(n):
                assert bool(t.sync_mask[ph] >> j & 1) == bool(tabs[j].sync_mask[ph] >> r & 1), "asymmetric barrier"
                assert bool(t.post_mask[ph] >> j & 1) == bool(tabs[j].post_mask[ph] >> r & 1), "asymmetric post"
            if flags & (UNI | ALLRANK | PAIRBAR):
                assert t.post_mask[ph] == 0  # the no-wait step exists only in the default bidirectional schedule
    if np_ and n > 1:
        assert all(t.sync_all[np_ - 1] == 1 for t in tabs)
    # ports[ph][r] = the ranks whose NVLink ports rank r's phase-ph transfer loads (both ends of its pair,
    # whoever issues); a transfer of phase ph + 1 may only start once every transfer of phase ph that shares
    # a port with it is over, i.e. its rank must be in the closing exchange of phase ph
    ports = [[set() for _ in range(n)] for _ in range(np_)]
    for ph in range(np_):
        for r, t in enumerate(tabs):
            if t.kind[0][ph] in (READ, WRITE, WARM) and t.peer[0][ph] != r:
                ports[ph][r] |= {r, t.peer[0][ph]}
                ports[ph][t.peer[0][ph]] |= {r, t.peer[0][ph]}
    for ph in range(np_ - 1):
        for x in range(n):
            for y in range(n):
                if x != y and ports[ph + 1][x] & ports[ph][y]:
                    if tabs[x].sync_mask[ph] >> y & 1:
                        continue
                    # the one exception: write -> read inside a round.  The ports are the PAIR's own on both sides of
                    # the barrier; y is only signalled (post), and whatever needs y's data waits for that signal itself
                    assert tabs[x].post_mask[ph] >> y & 1, f"phase {ph}: rank {x} may start while {y} still uses its port"
                    assert ports[ph][y] == ports[ph + 1][x] == {x, y}
                    assert tabs[x].kind[0][ph] in (WRITE, NONE) and tabs[x].kind[0][ph + 1] in (READ, NONE)
    reads, writes, verifies = {}, {}, {}
    for ph in range(np_):
        touched_by = {}
        for r, t in enumerate(tabs):
            k0, k1 = t.kind[0][ph], t.kind[1][ph]
            assert k1 in (NONE, VERIFY)
            if k0 in (READ, WRITE, WARM) and t.peer[0][ph] != r:
                p = t.peer[0][ph]
                assert 0 <= p < n
                assert p not in touched_by, "two ranks hit the same peer in one phase"
                touched_by[p] = r
                assert (t.sync_mask[ph] | t.post_mask[ph]) >> p & 1  # a pair always tells each other a transfer is over
                if flags & ALLRANK:
                    assert t.sync_all[ph] == 1  # round-1 behaviour: remote traffic closed by an all-rank barrier
                # the pairing is the tournament's: p's partner in that round is r
                assert any(plan.partner[rd][r] == p for rd in range(plan.rounds))
                if not (flags & UNI) and k0 != WARM:
                    assert tabs[p].kind[0][ph] == k0 and tabs[p].peer[0][ph] == r  # both ends issue at once
                if (flags & UNI) and k0 != WARM:
                    assert tabs[p].kind[0][ph] == NONE  # the partner is passive in this half
            if k0 == READ:
                key = (r, t.peer[0][ph])
                assert key not in reads
                reads[key] = ph
                if t.peer[0][ph] != r:
                    assert t.slot[0][ph] == slot_of(r, t.peer[0][ph])
            if k0 == WRITE:
                key = (r, t.peer[0][ph])
                assert key not in writes
                writes[key] = ph
                if t.peer[0][ph] != r:
                    assert t.slot[0][ph] == slot_of(r, t.peer[0][ph])
            for jb, k in ((0, k0), (1, k1)):
                if k == VERIFY:
                    key = (t.writer[jb][ph], r)  # (writer, owner)
                    assert t.peer[jb][ph] == r and key not in verifies
                    verifies[key] = (ph, t.slot[jb][ph])
            # CTA partitions
            if k0 != NONE:
                assert t.cta0[0][ph] == 0 and 0 < t.nctas[0][ph] <= ctas
            if k1 != NONE:
                assert t.cta0[1][ph] + t.nctas[1][ph] <= ctas and t.nctas[1][ph] >= 1
                if k0 != NONE and ctas >= 2:
                    assert t.cta0[0][ph] + t.nctas[0][ph] <= t.cta0[1][ph]
    pairs = {(i, j) for i in range(n) for j in range(n) if i != j}
    diag = {(i, i) for i in range(n)} if (n == 1 or flags & DIAG) else set()
    if ops & 1:
        assert set(reads) == pairs | diag
    else:
        assert not reads
    if ops & 2:
        assert set(writes) == pairs | diag
        assert set(verifies) == set(writes)  # every written slot is verified exactly once, by its owner
        for (w, o), (ph, slot) in verifies.items():
            assert ph > writes[(w, o)] or (w == o and ph > writes[(w, o)])
            if w != o:  # the writer signals the owner (and has published its checksum) at the write phase's own barrier ...
                wp = writes[(w, o)]
                assert (tabs[w].sync_mask[wp] | tabs[w].post_mask[wp]) >> o & 1
                # ... and the verify job waits for exactly that signal before it touches the slot
                jb = 0 if tabs[o].kind[0][ph] == VERIFY and tabs[o].writer[0][ph] == w else 1
                assert tabs[o].kind[jb][ph] == VERIFY and tabs[o].wait_barrier[jb][ph] == wp + 1
            else:
                jb = 0 if tabs[o].kind[0][ph] == VERIFY and tabs[o].writer[0][ph] == w else 1
                assert tabs[o].wait_barrier[jb][ph] == 0
            assert slot == (slot_of(w, o) if w != o else n - 1)
    else:
        assert not writes and not verifies


def test_default_8gpu_table_shape(pkg):
    """The headline configuration: warm-up, then 7 x (write, read + overlapped verify): 15 phases."""
    rc, tabs = table(pkg, 8)
    assert rc == 0 and tabs[0].n_phases == 15
    t = tabs[3]
    assert t.kind[0][0] == WARM
    assert [t.kind[0][p] for p in range(1, 15)] == [WRITE, READ] * 7
    assert [t.kind[1][p] for p in range(1, 15)] == [NONE, VERIFY] * 7
    assert all(t.nctas[0][p] == 148 - 32 and t.cta0[1][p] == 116 and t.nctas[1][p] == 32 for p in range(2, 15, 2))
    assert t.peer_mask == 0xFF & ~(1 << 3)
    # neighbourhood barriers: one all-rank exchange (the last); a write -> read barrier inside a round is the
    # pair alone; between rounds at most 4 ranks
    for t in tabs:
        assert [t.sync_all[p] for p in range(15)] == [0] * 14 + [1]
        assert bin(t.sync_mask[0]).count("1") == 1 and t.post_mask[0] == 0                      # warm -> W(0): the pair
        assert all(t.sync_mask[p] == 0 and bin(t.post_mask[p]).count("1") == 1 for p in range(1, 14, 2))  # W(r) -> R(r): no wait
        assert all(2 <= bin(t.sync_mask[p]).count("1") <= 4 and t.post_mask[p] == 0 for p in range(2, 14, 2))  # R(r) -> W(r+1)
    rc, pb = table(pkg, 8, flags=PAIRBAR)
    assert rc == 0 and all(bin(pb[0].sync_mask[p]).count("1") == 1 and pb[0].post_mask[p] == 0 for p in range(1, 14, 2))
    rc, old = table(pkg, 8, flags=ALLRANK)
    assert rc == 0 and all(old[0].sync_all[p] == 1 for p in range(15))


def test_overlap_does_not_depend_on_the_cta_count(pkg):
    """Every rank must walk the same number of phases whatever ITS grid size (a throttled rank next to full-size
    peers dead-locked the barrier sequence when it alone fell back to serial verify): only the split adapts."""
    shapes = {}
    for ctas in (148, 64, 8, 3, 2, 1):
        rc, tabs = table(pkg, 4, ctas=ctas, vctas=32)
        assert rc == 0
        t = tabs[1]
        shapes[ctas] = [(t.kind[0][p], t.kind[1][p], t.peer[0][p]) for p in range(t.n_phases)]
        for p in range(t.n_phases):
            if t.kind[1][p] == VERIFY and t.kind[0][p] != NONE:
                if ctas >= 64:
                    assert (t.nctas[0][p], t.cta0[1][p], t.nctas[1][p]) == (ctas - 32, ctas - 32, 32)
                elif ctas >= 2:
                    assert (t.nctas[0][p], t.cta0[1][p], t.nctas[1][p]) == (ctas - ctas // 2, ctas - ctas // 2, ctas // 2)
                else:
                    assert (t.cta0[0][p], t.nctas[0][p], t.cta0[1][p], t.nctas[1][p]) == (0, 1, 0, 1)  # one CTA, both jobs in turn
    assert all(s == shapes[148] for s in shapes.values())
    # mixed grids in one domain: rank 0 on 2 CTAs, the others on 148 — same phase count, same barrier masks
    rc0, s0 = schedule(pkg, 4, 0, ctas=2)
    rc1, s1 = schedule(pkg, 4, 1, ctas=148)
    assert rc0 == rc1 == 0 and s0.n_phases == s1.n_phases
    assert all(bool(s0.sync_mask[p] >> 1 & 1) == bool(s1.sync_mask[p] & 1) for p in range(s0.n_phases))
