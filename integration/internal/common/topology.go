// topology.go — new file for internal/common (NVIDIA/k8s-dra-driver-gpu): the NVML topology walk the
// fabric probe's caller and getCliqueID* share (SURVEY.md §8f n2).
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain).  Its semantics are pinned by the C++ twin
// k8s-dra-driver-gpu_b200/csrc/topo.cc (`cdprobe_topology`) and the oracle oracle/nvml_poll.c, which are
// checked against each other on every fake-NVML scenario (tests/test_oracle_nvml.py).
//
// It folds the two near-identical walks of cmd/compute-domain-kubelet-plugin/nvlib.go:208-363
// (getCliqueIDLegacy / getCliqueIDStrict) into one function and adds what the probe needs: per-GPU UUID,
// PCI bus id, MIG mode and active NVLink count.
package common

import (
	"fmt"

	"github.com/NVIDIA/go-nvml/pkg/nvml"
	"github.com/google/uuid"
)

const nvlinkMaxLinks = 18 // NVML_NVLINK_MAX_LINKS (vendor/.../go-nvml/pkg/nvml/const.go:47-48)

type GPUTopology struct {
	Index       int
	UUID        string
	PCIBusID    string
	MigEnabled  bool
	LinksActive int
	LinkMask    uint32 // bit l set: NVLink l is ENABLED
	FabricState uint8
}

type NodeTopology struct {
	GPUs []GPUTopology
	// CliqueID is "<clusterUUID>.<cliqueID>", or "" when the node is NVLink-capable but not
	// MNNVL-capable (zero cluster UUID) or fabric info is not supported.
	CliqueID string
}

// EnumerateTopology walks the devices the way VisitDevices does
// (vendor/github.com/NVIDIA/go-nvlib/pkg/nvlib/device/device.go:464-495).  strict selects the
// CrashOnNVLinkFabricErrors behaviour (nvlib.go:277-363): a fabric that is supported but not
// COMPLETED, or registered with an error status, is an error instead of "no clique".
// The caller owns NVML init/shutdown (nvlib.go:107-123).
func EnumerateTopology(lib nvml.Interface, strict bool) (*NodeTopology, error) {
	count, ret := lib.DeviceGetCount()
	if ret != nvml.SUCCESS {
		return nil, fmt.Errorf("error getting device count: %v", ret)
	}
	topo := &NodeTopology{}
	clusterUUIDs := map[string]struct{}{}
	cliqueIDs := map[string]struct{}{}
	for i := 0; i < count; i++ {
		dev, ret := lib.DeviceGetHandleByIndex(i)
		if ret != nvml.SUCCESS {
			return nil, fmt.Errorf("error getting device handle for index '%v': %v", i, ret)
		}
		g := GPUTopology{Index: i}
		if g.UUID, ret = dev.GetUUID(); ret != nvml.SUCCESS {
			return nil, fmt.Errorf("failed to read device uuid (%d): %v", i, ret)
		}
		if pci, ret := dev.GetPciInfo(); ret == nvml.SUCCESS {
			g.PCIBusID = busIDString(pci.BusId)
		}
		if cur, _, ret := dev.GetMigMode(); ret == nvml.SUCCESS {
			g.MigEnabled = cur == nvml.DEVICE_MIG_ENABLE
		}
		for l := 0; l < nvlinkMaxLinks; l++ {
			if st, ret := dev.GetNvLinkState(l); ret == nvml.SUCCESS && st == nvml.FEATURE_ENABLED {
				g.LinksActive++
				g.LinkMask |= 1 << uint(l)
			}
		}
		info, ret := dev.GetGpuFabricInfo()
		switch {
		case ret == nvml.ERROR_NOT_SUPPORTED: // no-clique fallback (nvlib.go:294-297)
		case ret != nvml.SUCCESS:
			return nil, fmt.Errorf("failed to get GPU fabric info (device %d/%s): %v", i, g.UUID, ret)
		default:
			g.FabricState = info.State
			attached, err := fabricAttached(i, g.UUID, info, strict)
			if err != nil {
				return nil, err
			}
			if attached {
				cu, err := uuid.FromBytes(info.ClusterUuid[:])
				if err != nil {
					return nil, fmt.Errorf("invalid cluster UUID (device %d/%s): %w", i, g.UUID, err)
				}
				clusterUUIDs[cu.String()] = struct{}{}
				cliqueIDs[fmt.Sprintf("%d", info.CliqueId)] = struct{}{}
			}
		}
		topo.GPUs = append(topo.GPUs, g)
	}
	if len(clusterUUIDs) == 0 && len(cliqueIDs) == 0 {
		return topo, nil
	}
	if len(clusterUUIDs) != 1 {
		return nil, fmt.Errorf("unexpected number of unique ClusterUUIDs found on devices")
	}
	if len(cliqueIDs) != 1 {
		return nil, fmt.Errorf("unexpected number of unique CliqueIDs found on devices")
	}
	for cu := range clusterUUIDs {
		for cq := range cliqueIDs {
			topo.CliqueID = fmt.Sprintf("%s.%s", cu, cq)
		}
	}
	return topo, nil
}

func fabricAttached(i int, duid string, info nvml.GpuFabricInfo, strict bool) (bool, error) {
	zero := info.ClusterUuid == [16]uint8{}
	if !strict { // IsFabricAttached, go-nvlib device.go:268-289
		return info.State == nvml.GPU_FABRIC_STATE_COMPLETED && !zero && nvml.Return(info.Status) == nvml.SUCCESS, nil
	}
	if info.State == nvml.GPU_FABRIC_STATE_NOT_SUPPORTED {
		return false, nil
	}
	if info.State != nvml.GPU_FABRIC_STATE_COMPLETED {
		return false, fmt.Errorf("NVLink fabric not attached (device %d/%s): state=%d, refusing to start", i, duid, info.State)
	}
	if nvml.Return(info.Status) != nvml.SUCCESS {
		return false, fmt.Errorf("NVLink fabric registration error (device %d/%s): status=%v, refusing to start", i, duid, nvml.Return(info.Status))
	}
	return !zero, nil
}

func busIDString(b [32]int8) string {
	out := make([]byte, 0, 32)
	for _, c := range b {
		if c == 0 {
			break
		}
		out = append(out, byte(c))
	}
	return string(out)
}
