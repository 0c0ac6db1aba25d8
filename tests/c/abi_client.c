/*
 * abi_client.c — a plain C caller of libcdprobe.so, the way a cgo preamble sees it: dlopen, resolve the
 * entry points of include/cdprobe.h, call them with caller-allocated structs.  No Python, no CUDA headers.
 *
 *   abi_client <libcdprobe.so> plan <n> <bytes> <mode>        host-only: prints the plan as JSON
 *   abi_client <libcdprobe.so> probe <bytes> <runs> [flags]   all visible GPUs: open, run x runs, close; JSON per run
 *
 * Exit codes: 0 ok, 3 probe not supported here (CDPROBE_ERR_NO_DEVICE / _UNSUPPORTED), 1 anything else.
 */
#include <dlfcn.h>
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cdprobe.h"

#define RESOLVE(var, name)                                  \
  *(void**)(&var) = dlsym(dl, name);                         \
  if (!var) {                                               \
    fprintf(stderr, "missing symbol %s\n", name);           \
    return 1;                                               \
  }

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <libcdprobe.so> plan|probe ...\n", argv[0]);
    return 1;
  }
  void* dl = dlopen(argv[1], RTLD_LAZY | RTLD_GLOBAL);
  if (!dl) {
    fprintf(stderr, "dlopen: %s\n", dlerror());
    return 1;
  }
  uint32_t (*abi_version)(void);
  const char* (*str_error)(int);
  const char* (*last_error)(void);
  int (*plan)(uint32_t, uint64_t, uint32_t, uint32_t, cdprobe_plan_t*);
  int (*open_)(const cdprobe_config_t*, cdprobe_t**);
  int (*run)(cdprobe_t*, cdprobe_result_t*);
  void (*close_)(cdprobe_t*);
  RESOLVE(abi_version, "cdprobe_abi_version")
  RESOLVE(str_error, "cdprobe_strerror")
  RESOLVE(last_error, "cdprobe_last_error")
  RESOLVE(plan, "cdprobe_plan")
  RESOLVE(open_, "cdprobe_open")
  RESOLVE(run, "cdprobe_run")
  RESOLVE(close_, "cdprobe_close")
  if (abi_version() != CDPROBE_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch\n");
    return 1;
  }
  if (strcmp(argv[2], "plan") == 0 && argc >= 6) {
    cdprobe_plan_t p;
    int rc = plan((uint32_t)atoi(argv[3]), strtoull(argv[4], NULL, 10), (uint32_t)atoi(argv[5]), 0, &p);
    if (rc != CDPROBE_OK) {
      fprintf(stderr, "cdprobe_plan: %s\n", str_error(rc));
      return 1;
    }
    printf("{\"n\": %u, \"rounds\": %u, \"bytes_per_pair\": %" PRIu64 ", \"partner\": [", p.n, p.rounds, p.bytes_per_pair);
    for (uint32_t r = 0; r < p.rounds; ++r) {
      printf("%s[", r ? ", " : "");
      for (uint32_t i = 0; i < p.n; ++i) printf("%s%d", i ? ", " : "", (int)p.partner[r][i]);
      printf("]");
    }
    printf("]}\n");
    return 0;
  }
  if (strcmp(argv[2], "probe") == 0 && argc >= 5) {
    cdprobe_config_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi = CDPROBE_ABI_VERSION;
    cfg.n_gpus = 0; /* all visible */
    cfg.bytes = strtoull(argv[3], NULL, 10);
    cfg.mode = CDPROBE_MODE_SLICED;
    cfg.flags = argc >= 6 ? (uint32_t)strtoul(argv[5], NULL, 0) : 0;
    cdprobe_t* h = NULL;
    int rc = open_(&cfg, &h);
    if (rc == CDPROBE_ERR_NO_DEVICE || rc == CDPROBE_ERR_UNSUPPORTED) {
      fprintf(stderr, "not supported: %s: %s\n", str_error(rc), last_error());
      return 3;
    }
    if (rc != CDPROBE_OK) {
      fprintf(stderr, "cdprobe_open: %s: %s\n", str_error(rc), last_error());
      return 1;
    }
    static cdprobe_result_t res; /* ~14 KB: caller-allocated, as in cgo */
    const int runs = atoi(argv[4]);
    for (int k = 0; k < runs; ++k) {
      rc = run(h, &res);
      if (rc != CDPROBE_OK) {
        fprintf(stderr, "cdprobe_run: %s: %s\n", str_error(rc), last_error());
        close_(h);
        return 1;
      }
      printf("{\"n\": %u, \"run_seq\": %" PRIu64 ", \"bytes_per_pair\": %" PRIu64 ", \"verdict\": %u, \"probe_ms\": %.4f, \"cells\": [",
             res.n, res.run_seq, res.bytes_per_pair, res.verdict, res.probe_ms);
      int first = 1;
      for (uint32_t i = 0; i < res.n; ++i)
        for (uint32_t j = 0; j < res.n; ++j) {
          const uint32_t c = i * CDPROBE_MAX_GPUS + j;
          printf("%s{\"i\": %u, \"j\": %u, \"rr\": %u, \"rw\": %u, \"sr\": \"%" PRIu64 "\", \"xr\": \"%" PRIu64
                 "\", \"sw\": \"%" PRIu64 "\", \"xw\": \"%" PRIu64 "\", \"gr\": %.1f, \"gw\": %.1f}",
                 first ? "" : ", ", i, j, res.reach_read[c], res.reach_write[c], res.sum_read[c], res.xor_read[c],
                 res.sum_write[c], res.xor_write[c], res.gbps_read[c], res.gbps_write[c]);
          first = 0;
        }
      printf("]}\n");
    }
    close_(h);
    return 0;
  }
  fprintf(stderr, "bad arguments\n");
  return 1;
}
