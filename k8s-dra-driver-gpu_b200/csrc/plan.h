// plan.h — tournament schedule and slice arithmetic of a probe (host only).
//
// ordered pairs (i, j), i != j, are partitioned by issuer
// i; round r pairs i with partner(i, r) (circle method) so every GPU has
// exactly one partner per round and owns both endpoints of the pair.
#pragma once
#include <stdint.h>

#include "../../include/cdprobe.h"
#include "probe_types.h"

namespace cdp {

struct Plan {
  uint32_t n = 0;          // ranks
  uint32_t rounds = 0;     // tournament rounds
  uint32_t n_slots = 0;    // landing slots per rank
  uint32_t n_slices = 0;   // source slices per rank
  uint32_t diag_slot = 0;  // slot/slice used by the loop-back (valid when diag)
  bool diag = false;
  bool full = false;
  uint64_t bpp = 0;        // bytes per pair
  uint64_t src_bytes = 0, land_bytes = 0;
  uint64_t src_off = 0, land_off = 0, alloc_bytes = 0;
  int8_t partner[kMaxRanks][kMaxRanks];  // [round][rank]
};

// Partner of rank i in round r of an n-rank tournament, -1 when i sits out (odd n).
int partner_of(uint32_t n, uint32_t r, uint32_t i);
// Slot of issuer i inside owner j's buffers (its index among j's peers).
inline uint32_t slot_of(uint32_t i, uint32_t j) { return i < j ? i : i - 1; }
// Returns CDPROBE_OK or CDPROBE_ERR_ARG.
int make_plan(uint32_t n, uint64_t bytes, uint32_t mode, uint32_t flags, Plan* out);

}  // namespace cdp
