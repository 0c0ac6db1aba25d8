/*
 * pattern.c — CPU restatement of the probe's integer definitions (SURVEY.md
 * §8d).  TEST INFRASTRUCTURE ONLY (see cdoracle.h); written independently of
 * the device code in k8s-dra-driver-gpu_b200/csrc/probe_types.h: scalar loops,
 * search instead of closed forms.
 *
 * The reference has no data pattern (no probe); the definitions
 * are frozen by SURVEY.md §8(d):
 *   source word   w[k] = splitmix64(seed ^ (rank << 56) ^ k)
 *   splitmix64(x): one step of Vigna's SplitMix64 with state x
 *                  (published vector: seed 1234567 -> 6457827717110365317, ...)
 *   write  word   w[k] = z ^ (z >> 32),  z = (salt + k) * 0x9E3779B97F4A7C15,
 *                 salt = splitmix64(seed ^ "WRITE" ^ (src << 56) ^ (dst << 48) ^ run_seq)
 *   checksum      S = sum of words mod 2^64;
 *                 X = xor over 16 KiB granules g of rotl64(xor of the granule's words, fold6(g))
 *   schedule      circle method: in round r ranks i, j meet when i + j == r (mod m),
 *                 m = n' - 1, n' = n rounded up to even; the rank with 2i == r meets n' - 1.
 */
#include <string.h>

#include "cdoracle.h"

#define GOLDEN 0x9E3779B97F4A7C15ull
#define GRANULE_WORDS 2048ull

uint64_t cdoracle_splitmix64(uint64_t x) {
  uint64_t z = x + GOLDEN;
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

uint64_t cdoracle_src_word(uint64_t seed, uint32_t rank, uint64_t k) {
  return cdoracle_splitmix64(seed ^ ((uint64_t)rank << 56) ^ k);
}

uint64_t cdoracle_write_salt(uint64_t seed, uint32_t src, uint32_t dst, uint64_t run_seq) {
  const uint64_t tag = 0x5752495445ull; /* "WRITE" */
  return cdoracle_splitmix64(seed ^ tag ^ ((uint64_t)src << 56) ^ ((uint64_t)dst << 48) ^ run_seq);
}

uint64_t cdoracle_write_word(uint64_t salt, uint64_t k) {
  const uint64_t z = (salt + k) * GOLDEN;
  return z ^ (z >> 32);
}

static uint32_t fold6(uint32_t g) {
  uint32_t f = 0;
  while (g) { /* xor of the 6-bit digits of g */
    f ^= g & 63u;
    g >>= 6;
  }
  return f;
}

static uint64_t rotl64(uint64_t x, uint32_t r) {
  while (r--) x = (x << 1) | (x >> 63);
  return x;
}

typedef struct {
  uint64_t sum, xr, gx, k;
} acc_t;

static void acc_word(acc_t* a, uint64_t w) {
  a->sum += w;
  a->gx ^= w;
  a->k++;
  if (a->k % GRANULE_WORDS == 0) {
    a->xr ^= rotl64(a->gx, fold6((uint32_t)(a->k / GRANULE_WORDS - 1)));
    a->gx = 0;
  }
}

static void acc_finish(acc_t* a, uint64_t* sum, uint64_t* xr) {
  if (a->k % GRANULE_WORDS != 0) a->xr ^= rotl64(a->gx, fold6((uint32_t)(a->k / GRANULE_WORDS)));
  *sum = a->sum;
  *xr = a->xr;
}

void cdoracle_checksum(const uint64_t* words, uint64_t n_words, uint64_t* sum, uint64_t* xr) {
  acc_t a = {0, 0, 0, 0};
  for (uint64_t k = 0; k < n_words; ++k) acc_word(&a, words[k]);
  acc_finish(&a, sum, xr);
}

void cdoracle_src_checksum(uint64_t seed, uint32_t rank, uint64_t first_word, uint64_t n_words, uint64_t* sum,
                           uint64_t* xr) {
  acc_t a = {0, 0, 0, 0};
  for (uint64_t k = 0; k < n_words; ++k) acc_word(&a, cdoracle_src_word(seed, rank, first_word + k));
  acc_finish(&a, sum, xr);
}

void cdoracle_write_checksum(uint64_t seed, uint32_t src, uint32_t dst, uint64_t run_seq, uint64_t n_words,
                             uint64_t* sum, uint64_t* xr) {
  const uint64_t salt = cdoracle_write_salt(seed, src, dst, run_seq);
  acc_t a = {0, 0, 0, 0};
  for (uint64_t k = 0; k < n_words; ++k) acc_word(&a, cdoracle_write_word(salt, k));
  acc_finish(&a, sum, xr);
}

uint32_t cdoracle_slot(uint32_t i, uint32_t j) { return i < j ? i : i - 1; }

int cdoracle_plan(uint32_t n, uint64_t bytes, uint32_t mode, uint32_t diag, cdoracle_plan_t* out) {
  if (n < 1 || n > CDORACLE_MAX_GPUS || out == 0 || mode > 2) return -1;
  memset(out, 0, sizeof(*out));
  memset(out->partner, -1, sizeof(out->partner));
  out->n = n;
  const uint32_t peers = n - 1;
  if (n == 1) diag = 1;
  uint64_t bpp;
  if (mode == 0) bpp = 65536;
  else if (mode == 1) {
    bpp = bytes / (peers ? peers : 1);
    bpp -= bpp % 128;
  } else {
    bpp = bytes - bytes % 128;
  }
  if (bpp < 128) return -1;
  out->bytes_per_pair = bpp;
  out->n_slots = peers + (diag ? 1 : 0);
  out->n_slices = mode == 2 ? 1 : out->n_slots;
  out->src_bytes = (uint64_t)out->n_slices * bpp;
  out->land_bytes = (uint64_t)out->n_slots * bpp;
  if (n == 1) return 0;
  const uint32_t ne = n + (n & 1u);
  const uint32_t m = ne - 1;
  out->rounds = m;
  for (uint32_t r = 0; r < m; ++r) {
    for (uint32_t i = 0; i < ne; ++i) {
      uint32_t p = ne; /* not found */
      if (i < m) {
        for (uint32_t j = 0; j < m; ++j)
          if (j != i && (i + j) % m == r) p = j;
        if (p == ne) p = m; /* 2i == r: meets the fixed rank */
      } else {
        for (uint32_t j = 0; j < m; ++j)
          if ((2 * j) % m == r) p = j;
      }
      if (i < n) out->partner[r][i] = (int8_t)(p < n ? (int)p : -1);
    }
  }
  return 0;
}
