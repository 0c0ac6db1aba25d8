// schedule.h — see schedule.cc.
#pragma once
#include <stdint.h>

#include "../../include/cdprobe.h"
#include "plan.h"
#include "probe_types.h"

namespace cdp {

struct ScheduleInput {
  const Plan* plan = nullptr;
  uint32_t rank = 0;         // global rank the table is for
  uint32_t ops = 0;          // CDPROBE_OP_*
  uint32_t flags = 0;        // CDPROBE_FLAG_* (OVERLAP_VERIFY, UNIDIRECTIONAL matter)
  uint32_t ctas = 0;         // CTAs of this rank's kernel
  uint32_t verify_ctas = 32; // CTAs of an overlapped verify job
  const int32_t (*status)[kMaxRanks] = nullptr;  // [issuer][owner] mapping status; null = every pair mapped
};

// Fills phases[0..kMaxPhases) / n_phases / peer_mask. CDPROBE_ERR_ARG when the table would overflow.
int make_phases(const ScheduleInput& in, Phase* phases, uint32_t* n_phases, uint32_t* peer_mask);

}  // namespace cdp
