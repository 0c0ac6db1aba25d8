/*
 * cdoracle.h — CPU oracle for the fabric probe.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library; libcdprobe.so (the product)
 * never links or calls it.
 *
 * PARITY UNPINNED: the reference (NVIDIA/k8s-dra-driver-gpu @ 2240711) holds
 * no golden vector, known-answer test or fixture for this path — it has no
 * NVLink probe at all and cmd/compute-domain-daemon has zero
 * tests — and it cannot be compiled here (Go toolchain absent; it needs
 * libnvidia-ml.so.1 and nvidia-imex-ctl at run time).  The algorithmic facts
 * live in closed-source third-party code: libnvidia-ml.so.1 reached through
 * github.com/NVIDIA/go-nvml v0.13.0-1.0.20260212130905-92cf8c963449
 * (go.mod:8) and github.com/NVIDIA/go-nvlib v0.10.0 (go.mod:7).  What this
 * oracle restates, with the call sites it follows:
 *
 *  (1) nvml_poll.c — the reference's CPU/NVML view of the node:
 *      init/shutdown discipline  cmd/compute-domain-kubelet-plugin/nvlib.go:107-123
 *      device walk               vendor/github.com/NVIDIA/go-nvlib/pkg/nvlib/device/device.go:464-495
 *      clique id (strict/legacy) cmd/compute-domain-kubelet-plugin/nvlib.go:208-363
 *      fabric-attached predicate vendor/.../go-nvlib/pkg/nvlib/device/device.go:268-310
 *      enumerate leg             cmd/gpu-kubelet-plugin/nvlib.go:457-531
 *      IMEX readiness gate       cmd/compute-domain-daemon/main.go:435-459
 *      NvLinkState / P2PStatus   vendor/.../go-nvml/pkg/nvml/device.go:1652-1661, :281-285
 *      and the frozen reachability definition of SURVEY.md §8(c).
 *  (2) pattern.c — the integer definitions of the probe's synthetic data
 *      (SURVEY.md §8d): splitmix64 source pattern (pinned against the published
 *      splitmix64 test vector, seed 1234567), write pattern, (S, X) checksums,
 *      round-robin tournament and slice arithmetic.
 */
#ifndef CDORACLE_H_
#define CDORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDORACLE_MAX_GPUS 16
#define CDORACLE_MAX_LINKS 18 /* NVML_NVLINK_MAX_LINKS, nvml.h:389 */

#define CDORACLE_FLAG_LEGACY_CLIQUE 0x1u /* getCliqueIDLegacy instead of Strict (gate CrashOnNVLinkFabricErrors off) */
#define CDORACLE_FLAG_NO_ENUMERATE 0x2u  /* skip the config-1 enumerate leg */
#define CDORACLE_FLAG_NO_IMEX_CTL 0x4u   /* do not exec nvidia-imex-ctl even when CLIQUE_ID is set */
#define CDORACLE_FLAG_THREADS 0x8u       /* one worker thread per GPU for the link and P2P polls */

typedef struct {
  uint32_t n;                                    /* GPUs NVML enumerates (after n_max clamp) */
  uint8_t reach[CDORACLE_MAX_GPUS * CDORACLE_MAX_GPUS]; /* [i*16+j], frozen definition SURVEY §8c */
  uint8_t link_active[CDORACLE_MAX_GPUS][CDORACLE_MAX_LINKS];
  uint8_t n_links[CDORACLE_MAX_GPUS];
  uint8_t mig_enabled[CDORACLE_MAX_GPUS];
  uint8_t fabric_state[CDORACLE_MAX_GPUS];
  int32_t fabric_ret[CDORACLE_MAX_GPUS];         /* nvmlReturn_t of GetGpuFabricInfo */
  int32_t fabric_status[CDORACLE_MAX_GPUS];
  uint32_t fabric_clique[CDORACLE_MAX_GPUS];
  uint8_t cluster_uuid[CDORACLE_MAX_GPUS][16];
  int32_t p2p_read[CDORACLE_MAX_GPUS * CDORACLE_MAX_GPUS];   /* nvmlGpuP2PStatus_t, or -(nvmlReturn_t) */
  int32_t p2p_write[CDORACLE_MAX_GPUS * CDORACLE_MAX_GPUS];
  int32_t p2p_nvlink[CDORACLE_MAX_GPUS * CDORACLE_MAX_GPUS];
  char uuid[CDORACLE_MAX_GPUS][96];
  char name[CDORACLE_MAX_GPUS][96];
  char pci_bus_id[CDORACLE_MAX_GPUS][32];
  uint64_t memory_total[CDORACLE_MAX_GPUS];
  int32_t minor[CDORACLE_MAX_GPUS];
  int32_t cc_major[CDORACLE_MAX_GPUS], cc_minor[CDORACLE_MAX_GPUS];
  char driver_version[96];
  int32_t cuda_driver_version;
  char clique_id[96];                            /* "<clusterUUID>.<cliqueId>" or "" */
  int32_t clique_err;                            /* != 0: getCliqueID would return an error */
  char clique_err_text[160];
  int32_t imex_gate;                             /* -1 not applicable (clique ""), 0 not ready, 1 READY */
  uint32_t nvml_calls;
  double init_ms, enumerate_ms, fabric_ms, link_poll_ms, p2p_poll_ms, imex_ms, shutdown_ms, total_ms;
} cdoracle_nvml_t;

/* 0 ok; -1 libnvidia-ml.so.1 cannot be loaded; -2 a symbol is missing; >0 nvmlReturn_t of a failed call. */
int cdoracle_nvml_poll(uint32_t n_max, uint32_t flags, cdoracle_nvml_t* out);

/* ---- pattern.c ----------------------------------------------------------------------- */
uint64_t cdoracle_splitmix64(uint64_t x);
uint64_t cdoracle_src_word(uint64_t seed, uint32_t rank, uint64_t k);
uint64_t cdoracle_write_salt(uint64_t seed, uint32_t src, uint32_t dst, uint64_t run_seq);
uint64_t cdoracle_write_word(uint64_t salt, uint64_t k);
/* (S, X) of n_words 64-bit little-endian words. */
void cdoracle_checksum(const uint64_t* words, uint64_t n_words, uint64_t* sum, uint64_t* xr);
/* (S, X) of words [first_word, first_word + n_words) of rank's source buffer, generated on the fly. */
void cdoracle_src_checksum(uint64_t seed, uint32_t rank, uint64_t first_word, uint64_t n_words, uint64_t* sum, uint64_t* xr);
/* (S, X) of the n_words a writer src stores into dst's landing slot in run run_seq. */
void cdoracle_write_checksum(uint64_t seed, uint32_t src, uint32_t dst, uint64_t run_seq, uint64_t n_words, uint64_t* sum,
                             uint64_t* xr);

typedef struct {
  uint32_t n, rounds, n_slots, n_slices;
  uint64_t bytes_per_pair, src_bytes, land_bytes;
  int8_t partner[CDORACLE_MAX_GPUS][CDORACLE_MAX_GPUS]; /* [round][rank] */
} cdoracle_plan_t;
/* mode: 0 reach-only, 1 sliced, 2 full; diag: loop-back slot wanted. 0 ok, -1 bad argument. */
int cdoracle_plan(uint32_t n, uint64_t bytes, uint32_t mode, uint32_t diag, cdoracle_plan_t* out);
/* Which slot/slice of owner j issuer i uses. */
uint32_t cdoracle_slot(uint32_t i, uint32_t j);

#ifdef __cplusplus
}
#endif
#endif
