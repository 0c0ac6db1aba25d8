"""The C oracle (oracle/pattern.c) against the golden vectors (tests/golden/golden.json)."""
import ctypes as C

import pytest


def test_splitmix64_published_vector(oracle, golden):
    # Vigna's reference SplitMix64, seed 1234567: the one external known-answer vector we have
    s = 1234567
    got = []
    for _ in range(5):
        got.append(oracle.lib().cdoracle_splitmix64(s))
        s = (s + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    assert [str(g) for g in got] == golden["splitmix64_seed_1234567"]
    assert got[0] == 6457827717110365317


def test_src_and_write_words(oracle, golden):
    L = oracle.lib()
    for v in golden["src_words"]:
        assert L.cdoracle_src_word(int(v["seed"]), v["rank"], v["k"]) == int(v["word"])
    for v in golden["write_words"]:
        salt = L.cdoracle_write_salt(int(v["seed"]), v["src"], v["dst"], v["run_seq"])
        assert salt == int(v["salt"])
        assert L.cdoracle_write_word(salt, v["k"]) == int(v["word"])


def test_checksums_match_golden(oracle, golden):
    for v in golden["src_checksums"]:
        s, x = oracle.src_checksum(int(v["seed"]), v["rank"], v["first_word"], v["n_words"])
        assert (str(s), str(x)) == (v["sum"], v["xor"])
    for v in golden["write_checksums"]:
        s, x = oracle.write_checksum(int(v["seed"]), v["src"], v["dst"], v["run_seq"], v["n_words"])
        assert (str(s), str(x)) == (v["sum"], v["xor"])


def test_buffer_checksum_equals_streaming(oracle):
    L = oracle.lib()
    for n in (0, 1, 2047, 2048, 2049, 5000):
        words = (C.c_uint64 * max(n, 1))()
        for k in range(n):
            words[k] = L.cdoracle_src_word(oracle.DEFAULT_SEED, 2, 100 + k)
        s, x = C.c_uint64(), C.c_uint64()
        L.cdoracle_checksum(words, n, C.byref(s), C.byref(x))
        assert (s.value, x.value) == oracle.src_checksum(oracle.DEFAULT_SEED, 2, 100, n)


def test_checksum_is_position_sensitive_across_granules(oracle):
    # swapping two 16 KiB granules keeps S but changes X (a misplaced chunk is detected)
    L = oracle.lib()
    n = 2048 * 3
    words = (C.c_uint64 * n)()
    for k in range(n):
        words[k] = L.cdoracle_src_word(oracle.DEFAULT_SEED, 0, k)
    s0, x0 = C.c_uint64(), C.c_uint64()
    L.cdoracle_checksum(words, n, C.byref(s0), C.byref(x0))
    for k in range(2048):
        words[k], words[2048 + k] = words[2048 + k], words[k]
    s1, x1 = C.c_uint64(), C.c_uint64()
    L.cdoracle_checksum(words, n, C.byref(s1), C.byref(x1))
    assert s0.value == s1.value and x0.value != x1.value


def test_plans_match_golden(oracle, golden):
    for g in golden["plans"]:
        p = oracle.plan(g["n"], g["bytes"], g["mode"], g["diag"])
        assert p.bytes_per_pair == g["bytes_per_pair"]
        assert (p.n_slots, p.n_slices, p.rounds) == (g["n_slots"], g["n_slices"], g["rounds"])
        got = [[p.partner[r][i] for i in range(g["n"])] for r in range(g["rounds"])]
        assert got == g["partner"]


@pytest.mark.parametrize("n", list(range(1, 17)))
def test_schedule_is_a_one_factorisation(oracle, n):
    """Every unordered pair meets exactly once; nobody has two partners in a round."""
    p = oracle.plan(n, 1 << 30, 1)
    met = set()
    for r in range(p.rounds):
        row = [p.partner[r][i] for i in range(n)]
        for i, j in enumerate(row):
            if j < 0:
                assert n % 2 == 1
                continue
            assert j != i and row[j] == i
            if i < j:
                assert (i, j) not in met
                met.add((i, j))
        assert sum(1 for j in row if j < 0) == (n % 2 if n > 1 else 0)
    assert len(met) == n * (n - 1) // 2
    assert p.rounds == (0 if n == 1 else (n if n % 2 else n - 1))


def test_headline_bytes_per_pair(oracle):
    # N=8, B=1 GiB sliced -> 153 391 616 B per pair, 1 073 741 312 B per GPU
    p = oracle.plan(8, 1 << 30, 1)
    assert p.bytes_per_pair == 153391616
    assert 7 * p.bytes_per_pair == 1073741312
    assert oracle.plan(2, 64 << 20, 2).bytes_per_pair == 64 << 20
    assert oracle.plan(8, 1 << 30, 0).bytes_per_pair == 65536
