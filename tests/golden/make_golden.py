#!/usr/bin/env python3
"""Generates tests/golden/golden.json — known-answer vectors for the probe's integer definitions.

The reference (NVIDIA/k8s-dra-driver-gpu) holds no fixture for this path (SURVEY.md §8c:
"parity unpinned"), and being Go it cannot be imported here, so these vectors come from a
third, pure-Python statement of SURVEY.md §8(d) written independently of both the C oracle
(oracle/pattern.c) and the device code.  One vector is external: the published SplitMix64
test vector (Vigna's reference implementation, seed 1234567).

Run:  python tests/golden/make_golden.py   (rewrites golden.json deterministically)
"""
import json
import os

M = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15
SEED = 0xCD5EED0000000001
GRANULE_WORDS = 16384 // 8


def splitmix64(x):
    z = (x + GOLDEN) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def src_word(seed, rank, k):
    return splitmix64(seed ^ (rank << 56) ^ k)


def write_salt(seed, src, dst, run_seq):
    return splitmix64(seed ^ int.from_bytes(b"WRITE", "big") ^ (src << 56) ^ (dst << 48) ^ run_seq)


def write_word(salt, k):
    z = ((salt + k) * GOLDEN) & M
    return z ^ (z >> 32)


def rotl(x, r):
    r %= 64
    return ((x << r) | (x >> (64 - r))) & M if r else x


def fold6(g):
    f = 0
    while g:
        f ^= g & 63
        g >>= 6
    return f


def checksum(words):
    s = 0
    x = 0
    for g in range(0, len(words), GRANULE_WORDS):
        gx = 0
        for w in words[g:g + GRANULE_WORDS]:
            s = (s + w) & M
            gx ^= w
        x ^= rotl(gx, fold6(g // GRANULE_WORDS))
    return s, x


def partner(n, r, i):
    ne = n + (n & 1)
    m = ne - 1
    if i == ne - 1:
        p = next(x for x in range(m) if (2 * x) % m == r)
    else:
        j = (r - i) % m
        p = ne - 1 if j == i else j
    return p if p < n else -1


def plan(n, nbytes, mode, diag=False):
    peers = n - 1
    diag = diag or n == 1
    if mode == 0:
        bpp = 65536
    elif mode == 1:
        bpp = nbytes // max(peers, 1) // 128 * 128
    else:
        bpp = nbytes // 128 * 128
    n_slots = peers + (1 if diag else 0)
    rounds = 0 if n == 1 else (n if n & 1 else n - 1)
    return {
        "n": n, "bytes": nbytes, "mode": mode, "diag": diag, "bytes_per_pair": bpp,
        "n_slots": n_slots, "n_slices": 1 if mode == 2 else n_slots, "rounds": rounds,
        "partner": [[partner(n, r, i) for i in range(n)] for r in range(rounds)],
    }


def main():
    out = {}
    # external vector: SplitMix64 reference implementation, seed 1234567, first five outputs
    s = 1234567
    vec = []
    for _ in range(5):
        vec.append(splitmix64(s))
        s = (s + GOLDEN) & M
    assert vec[0] == 6457827717110365317 and vec[4] == 16408922859458223821
    out["splitmix64_seed_1234567"] = [str(v) for v in vec]

    out["src_words"] = [
        {"seed": str(SEED), "rank": r, "k": k, "word": str(src_word(SEED, r, k))}
        for r in (0, 1, 7, 15) for k in (0, 1, 2047, 2048, (1 << 27) - 1)
    ]
    out["write_words"] = [
        {"seed": str(SEED), "src": a, "dst": b, "run_seq": q, "k": k,
         "salt": str(write_salt(SEED, a, b, q)), "word": str(write_word(write_salt(SEED, a, b, q), k))}
        for (a, b, q) in ((0, 1, 2), (1, 0, 2), (7, 3, 1000)) for k in (0, 1, 1023, 1 << 20)
    ]
    cks = []
    for rank, first, n_words in ((0, 0, 16), (0, 0, 2048), (3, 8192, 2048 * 3 + 16), (7, 1 << 20, 2048 * 65 + 1008),
                                 (1, 0, 8192)):
        words = [src_word(SEED, rank, first + k) for k in range(n_words)]
        s_, x_ = checksum(words)
        cks.append({"seed": str(SEED), "rank": rank, "first_word": first, "n_words": n_words,
                    "sum": str(s_), "xor": str(x_)})
    out["src_checksums"] = cks
    wck = []
    for (a, b, q, n_words) in ((0, 1, 2, 8192), (5, 2, 9, 2048 * 4 + 16)):
        salt = write_salt(SEED, a, b, q)
        s_, x_ = checksum([write_word(salt, k) for k in range(n_words)])
        wck.append({"seed": str(SEED), "src": a, "dst": b, "run_seq": q, "n_words": n_words,
                    "sum": str(s_), "xor": str(x_)})
    out["write_checksums"] = wck
    out["plans"] = [plan(n, b, m) for (n, b, m) in
                    ((1, 1 << 30, 1), (2, 64 << 20, 2), (2, 1 << 30, 1), (3, 1 << 30, 1), (4, 1 << 30, 1),
                     (8, 1 << 30, 1), (8, 1 << 30, 2), (8, 1 << 30, 0), (5, 1000003, 1), (16, 1 << 30, 1))]
    # SURVEY.md §8(d) quotes these for N=8, B=1 GiB, sliced
    p8 = plan(8, 1 << 30, 1)
    assert p8["bytes_per_pair"] == 153391616 and 7 * p8["bytes_per_pair"] == 1073741312
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print(path)


if __name__ == "__main__":
    main()
