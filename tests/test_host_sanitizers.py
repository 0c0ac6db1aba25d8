"""ASAN + UBSAN over the product's host-only logic (plan, schedule, rendezvous, topology): the four
.cc files that need no CUDA are compiled straight into a sanitised binary and driven through the C ABI
across the whole argument space (the CUDA side is covered by compute-sanitizer on the GPU box)."""
import os
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "k8s-dra-driver-gpu_b200", "csrc")
FAKE = os.path.join(ROOT, "tests", "fake_nvml", "libnvidia-ml.so.1")


@pytest.fixture(scope="module")
def binary(oracle, tmp_path_factory):
    exe = tmp_path_factory.mktemp("asan") / "host_sanitize"
    srcs = [os.path.join(ROOT, "tests", "c", "host_sanitize.cc")] + [os.path.join(CSRC, f) for f in
                                                                     ("plan.cc", "schedule.cc", "rendezvous.cc", "topo.cc")]
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", "-I/usr/local/cuda/include", "-o", str(exe), *srcs, "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("libasan/libubsan not installed for this g++")
    assert r.returncode == 0, r.stderr[-2000:]
    return str(exe)


@pytest.mark.parametrize("scenario", ["gpus 8\n", "gpus 16\nmig 3 1\nlink_down 2 4\n",
                                      "gpus 4\nfabric_all 3 0 7 00112233445566778899aabbccddeeff\nfabric 2 2 0 7 00112233445566778899aabbccddeeff\n"])
def test_host_logic_is_clean_under_asan_ubsan(binary, tmp_path, scenario):
    sc = tmp_path / "scenario.txt"
    sc.write_text(scenario)
    env = dict(os.environ, CDPROBE_NVML_PATH=FAKE, FAKE_NVML_SCENARIO=str(sc),
               ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([binary], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "HOST_SANITIZE_DONE" in r.stdout
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
