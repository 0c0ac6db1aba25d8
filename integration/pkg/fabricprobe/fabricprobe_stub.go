//go:build !cgo

// Stub so that `go build ./...` keeps working with CGO_ENABLED=0 (the reference's Makefile
// builds with CGO_ENABLED=1, Makefile:56-59, but tooling such as golangci-lint may not).
package fabricprobe

import (
	"context"
	"errors"
)

var (
	ErrUnsupported = errors.New("fabricprobe: not supported on this node")
	ErrTimeout     = errors.New("fabricprobe: probe timed out")
	ErrState       = errors.New("fabricprobe: handle is unusable")
	ErrCUDA        = errors.New("fabricprobe: CUDA call failed")
)

const (
	ModeReachOnly, ModeSliced, ModeFull             = 0, 1, 2
	OpRead, OpWrite                                 = 1, 2
	FlagFabricHandles, FlagMigAware, FlagLocalDiag = 0x01, 0x02, 0x04
)

type Config struct {
	LibraryPath string
	Ordinals    []int
	Bytes       uint64
	Mode        uint32
	Ops         uint32
	TimeoutMs   uint32
	Flags       uint32
	MinFraction  float32
	LinkPeakGBps float32
}

type Result struct {
	N                     int
	ReachRead, ReachWrite []bool
	GBpsRead, GBpsWrite   []float32
	Status                []int32
	ProbeMs               float64
	Verdict, Aborted      bool
	BytesPerPair          uint64
	MinGBpsRead, MinGBpsWrite   float32
	GateGBpsRead, GateGBpsWrite float32
	UnreachablePairs, SlowPairs int
}

type Probe struct{}

func Open(Config) (*Probe, error)                  { return nil, ErrUnsupported }
func (*Probe) Run(context.Context) (Result, error) { return Result{}, ErrUnsupported }
func (*Probe) Close()                              {}
