// fabricprobe.go — new file for cmd/compute-domain-daemon (NVIDIA/k8s-dra-driver-gpu).
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image, SURVEY.md F4).  The behaviour
// it specifies is executable as the C++ mirror k8s-dra-driver-gpu_b200/csrc/daemon_main.cc
// (`cdprobe-daemon {run,check}`, tests/test_daemon.py).
//
// Wiring (two call sites in main.go, see INTEGRATION.md §2):
//   run():   after addComputeDomainCliqueLabel(), BEFORE the `if flags.cliqueID == ""` early wait
//            (main.go:244-250) so single-node HGX boxes are covered too:
//                stopProbe := startFabricProbe(ctx, flags, controllerUpdates)
//                defer stopProbe()
//   check(): after the existing IMEX gate (main.go:435-459):
//                if err := checkFabricProbeVerdict(); err != nil { return err }
package main

import (
	"context"
	"encoding/json"
	"errors"
	"fmt"
	"os"
	"path/filepath"
	"time"

	"k8s.io/klog/v2"

	"sigs.k8s.io/dra-driver-nvidia-gpu/pkg/fabricprobe"
	"sigs.k8s.io/dra-driver-nvidia-gpu/pkg/featuregates"
)

const (
	// The per-ComputeDomain bind mount shared by `run` and `check` (computedomain.go:170-177).
	fabricProbeVerdictPath = "/imexd/fabricprobe.json"
)

type fabricProbeVerdict struct {
	TimeUnix         int64     `json:"time_unix"`
	OK               bool      `json:"ok"`
	N                int       `json:"n"`
	UnreachablePairs int       `json:"unreachable_pairs"`
	MinGBpsRead      float32   `json:"min_gbps_read"`
	MinGBpsWrite     float32   `json:"min_gbps_write"`
	ProbeMs          float64   `json:"probe_ms"`
	ReachRead        []bool    `json:"reach_read"`
	ReachWrite       []bool    `json:"reach_write"`
	GBpsRead         []float32 `json:"gbps_read"`
	GBpsWrite        []float32 `json:"gbps_write"`
	Error            string    `json:"error"`
}

// startFabricProbe opens the probe once and re-runs it whenever the set of daemons in the domain
// changes (the same signal that drives IMEXDaemonUpdateLoopWithDNSNames, main.go:384-431).
func startFabricProbe(ctx context.Context, flags *Flags, updates <-chan struct{}) (stop func()) {
	stop = func() {}
	if !featuregates.Enabled(featuregates.FabricProbe) {
		return stop
	}
	probe, err := fabricprobe.Open(fabricprobe.Config{
		Bytes:       flags.fabricProbeBytes,       // FABRIC_PROBE_BYTES, default 1 GiB
		Mode:        flags.fabricProbeMode,        // FABRIC_PROBE_MODE, default sliced
		MinFraction: flags.fabricProbeMinFraction, // FABRIC_PROBE_MIN_FRACTION, 0 = library default
		Flags:       fabricprobe.FlagFabricHandles | fabricprobe.FlagMigAware,
	})
	switch {
	case errors.Is(err, fabricprobe.ErrUnsupported):
		// No libcdprobe.so / no CUDA driver / not sm_100: there is no CPU stand-in. No verdict is
		// written and check() does not gate on a missing verdict.
		klog.Infof("fabric probe not supported on this node: %v", err)
		return stop
	case err != nil:
		klog.Errorf("error opening fabric probe: %v", err)
		return stop
	}
	runOnce := func() {
		t0 := time.Now()
		res, err := probe.Run(ctx)
		klog.V(6).Infof("t_fabric_probe %.6f s", time.Since(t0).Seconds())
		v := fabricProbeVerdict{TimeUnix: time.Now().Unix(), OK: err == nil && res.Verdict, N: res.N,
			ProbeMs: res.ProbeMs, ReachRead: res.ReachRead, ReachWrite: res.ReachWrite,
			GBpsRead: res.GBpsRead, GBpsWrite: res.GBpsWrite}
		if err != nil {
			v.Error = err.Error()
		}
		for i := 0; i < res.N; i++ {
			for j := 0; j < res.N; j++ {
				if i != j && !(res.ReachRead[i*res.N+j] && res.ReachWrite[i*res.N+j]) {
					v.UnreachablePairs++
				}
			}
		}
		if err := writeFileAtomic(fabricProbeVerdictPath, v); err != nil {
			klog.Errorf("cannot write %s: %v", fabricProbeVerdictPath, err)
		}
		klog.Infof("fabric probe: verdict ok=%t, %d GPU(s), %d unreachable pair(s), %.3f ms", v.OK, v.N,
			v.UnreachablePairs, v.ProbeMs)
	}
	done := make(chan struct{})
	go func() {
		defer close(done)
		runOnce()
		for {
			select {
			case <-ctx.Done():
				return
			case _, ok := <-updates:
				if !ok {
					return
				}
				runOnce()
			}
		}
	}()
	return func() {
		<-done
		probe.Close()
	}
}

// checkFabricProbeVerdict is the addition to check(): a failed verdict makes the pod NotReady; a missing
// one does not (same spirit as the reference's no-op when CLIQUE_ID is empty, main.go:436-439).
func checkFabricProbeVerdict() error {
	if !featuregates.Enabled(featuregates.FabricProbe) {
		return nil
	}
	raw, err := os.ReadFile(fabricProbeVerdictPath)
	if errors.Is(err, os.ErrNotExist) {
		return nil
	}
	if err != nil {
		return fmt.Errorf("fabric probe verdict unreadable: %w", err)
	}
	var v fabricProbeVerdict
	if err := json.Unmarshal(raw, &v); err != nil {
		return fmt.Errorf("fabric probe verdict unreadable: %w", err)
	}
	if !v.OK {
		return fmt.Errorf("fabric probe failed: %d unreachable pair(s), min read %.0f GB/s, min write %.0f GB/s: %s",
			v.UnreachablePairs, v.MinGBpsRead, v.MinGBpsWrite, v.Error)
	}
	return nil
}

func writeFileAtomic(path string, v any) error {
	raw, err := json.MarshalIndent(v, "", " ")
	if err != nil {
		return err
	}
	tmp := filepath.Join(filepath.Dir(path), "."+filepath.Base(path)+".tmp")
	if err := os.WriteFile(tmp, raw, 0o644); err != nil {
		return err
	}
	return os.Rename(tmp, path)
}
