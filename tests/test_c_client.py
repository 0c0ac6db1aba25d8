"""A plain C program (tests/c/abi_client.c: dlopen + include/cdprobe.h, nothing else) drives the ABI the
way the cgo shim would: compiled-language caller, caller-allocated structs, integer error codes."""
import json
import os
import subprocess

import pytest

from conftest import ROOT, gpu_count

SRC = os.path.join(ROOT, "tests", "c", "abi_client.c")
SEED = 0xCD5EED0000000001


@pytest.fixture(scope="module")
def client(pkg, tmp_path_factory):
    exe = tmp_path_factory.mktemp("c") / "abi_client"
    subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-std=c11", "-o", str(exe), SRC, "-ldl"], check=True)
    return str(exe), pkg.abi.LIB_PATH


def test_c_client_plan_matches_golden(client, golden):
    exe, lib = client
    for g in golden["plans"]:
        if g["diag"] and g["n"] > 1:
            continue
        out = subprocess.run([exe, lib, "plan", str(g["n"]), str(g["bytes"]), str(g["mode"])], capture_output=True,
                             text=True, check=True).stdout
        j = json.loads(out)
        assert j["bytes_per_pair"] == g["bytes_per_pair"] and j["rounds"] == g["rounds"] and j["partner"] == g["partner"]


@pytest.mark.skipif(gpu_count() > 0, reason="CPU-only behaviour")
def test_c_client_probe_fails_loudly_without_gpu(client):
    exe, lib = client
    r = subprocess.run([exe, lib, "probe", str(1 << 20), "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_client_probe_parity(client, oracle):
    """All visible GPUs, driven from C: every cell's reachability and checksums equal the oracle's."""
    exe, lib = client
    nbytes = 8 << 20
    r = subprocess.run([exe, lib, "probe", str(nbytes), "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    runs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(runs) == 2
    for run in runs:
        n = run["n"]
        assert n == gpu_count() and run["bytes_per_pair"] == oracle.plan(n, nbytes, 1).bytes_per_pair
        for c in run["cells"]:
            i, j = c["i"], c["j"]
            if i == j and n > 1:
                assert c["rr"] == 1 and c["rw"] == 1
                continue
            assert c["rr"] == 1 and c["rw"] == 1
            assert (int(c["sr"]), int(c["xr"])) == oracle.expected_read(SEED, n, nbytes, 1, i, j)
            assert (int(c["sw"]), int(c["xw"])) == oracle.write_checksum(SEED, i, j, run["run_seq"], run["bytes_per_pair"] // 8)
