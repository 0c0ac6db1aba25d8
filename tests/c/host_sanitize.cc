// host_sanitize.cc — drives the host-only entry points of the product (plan, schedule, rendezvous, topology)
// from a binary built with -fsanitize=address,undefined (tests/test_host_sanitizers.py).  The CUDA-facing
// files are covered on the GPU box by compute-sanitizer; this covers the pure host logic.
#include <stdio.h>

#include <initializer_list>
#include <string.h>
#include <unistd.h>

#include "../../include/cdprobe.h"

int main(int argc, char** argv) {
  int checked = 0;
  for (uint32_t n = 0; n <= 17; ++n)
    for (uint32_t mode = 0; mode <= 3; ++mode)
      for (uint32_t flags : {0u, 0x04u}) {
        cdprobe_plan_t p;
        const uint64_t sizes[] = {0, 127, 128, 1000003, 1ull << 30, 1ull << 40};
        for (uint64_t b : sizes) {
          int rc = cdprobe_plan(n, b, mode, flags, &p);
          if (rc == CDPROBE_OK && (p.bytes_per_pair % 128 != 0 || p.n != n)) return 2;
          checked++;
        }
      }
  for (uint32_t n = 1; n <= 16; ++n)
    for (uint32_t rank = 0; rank < n; ++rank)
      for (uint32_t ops = 0; ops <= 3; ++ops)
        for (uint32_t flags : {0u, 0x80u, 0x100u, 0x180u, 0x04u, 0x84u})
          for (uint32_t ctas : {1u, 2u, 8u, 63u, 64u, 148u, 65535u}) {
            cdprobe_schedule_t s;
            int rc = cdprobe_schedule(n, rank, 1ull << 30, CDPROBE_MODE_SLICED, ops, flags, ctas, 32, &s);
            if (rc == CDPROBE_OK && s.n_phases > CDPROBE_MAX_PHASES) return 3;
            checked++;
          }
  cdprobe_topology_t t;
  int rc = cdprobe_topology(1, &t);
  printf("topology rc=%d n=%u clique=\"%s\" err=\"%s\"\n", rc, t.n, t.clique_id, t.clique_error);
  rc = cdprobe_topology(0, &t);
  char session[64];
  snprintf(session, sizeof(session), "asan-%d", (int)getpid());
  rc = cdprobe_rendezvous_selftest(session, 0, 1, 1000);
  if (rc != CDPROBE_OK) return 4;
  if (cdprobe_rendezvous_selftest(session, 0, 2, 50) != CDPROBE_ERR_RENDEZVOUS) return 5;  // nobody else: clean timeout
  printf("HOST_SANITIZE_DONE %d\n", checked);
  (void)argc;
  (void)argv;
  return 0;
}
