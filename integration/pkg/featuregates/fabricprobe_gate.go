// fabricprobe_gate.go — addition to pkg/featuregates (NVIDIA/k8s-dra-driver-gpu).
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain).  In the reference tree these two declarations go
// INTO featuregates.go — the constant into the const block (featuregates.go:47-81), the spec into
// defaultFeatureGates (featuregates.go:92-151); they are kept in a file of their own here so that the
// patch is reviewable without a diff tool.  A map literal cannot be extended from a second file, hence
// the init(): the package-level singleton is only built on first use (sync.Once, featuregates.go:153-160),
// after every init() has run.
package featuregates

import (
	"k8s.io/apimachinery/pkg/util/version"
	"k8s.io/component-base/featuregate"
)

const (
	// FabricProbe makes the compute-domain-daemon run the all-pairs NVLink reachability + bandwidth
	// probe (pkg/fabricprobe -> libcdprobe.so) on the GPUs it owns and lets `check` gate the pod's
	// readiness on the probe's verdict, next to the IMEX daemon's READY (cmd/compute-domain-daemon).
	FabricProbe featuregate.Feature = "FabricProbe"
)

func init() {
	defaultFeatureGates[FabricProbe] = featuregate.VersionedSpecs{
		{
			Default:    false,
			PreRelease: featuregate.Alpha,
			Version:    version.MajorMinor(0, 4),
		},
	}
}
