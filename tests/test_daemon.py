"""`cdprobe-daemon {run,check}` — the fabric-probe slice of cmd/compute-domain-daemon.

`check` must keep the reference's behaviour (cmd/compute-domain-daemon/main.go:435-459): no-op text
when CLIQUE_ID is empty, otherwise `nvidia-imex-ctl -c /imexd/imexd.cfg -q` == "READY\\n"; on top of
that it consults the verdict file `run` writes.  The reference has no unit test for the daemon (SURVEY
§4), so these are written the way its bats tests assert: on exit codes and output text."""
import json
import os
import subprocess

import pytest

from conftest import ROOT, gpu_count

DAEMON = os.path.join(ROOT, "k8s-dra-driver-gpu_b200", "cdprobe-daemon")
LIB = os.path.join(ROOT, "k8s-dra-driver-gpu_b200", "libcdprobe.so")


def daemon(args, env=None, timeout=120):
    e = {"PATH": os.environ.get("PATH", ""), "LD_LIBRARY_PATH": os.environ.get("LD_LIBRARY_PATH", "")}
    e.update(env or {})
    return subprocess.run([DAEMON, *args], env=e, capture_output=True, text=True, timeout=timeout)


def fake_ctl(tmp_path, text, code=0):
    p = tmp_path / "nvidia-imex-ctl"
    p.write_text(f"#!/bin/sh\n[ \"$1\" = -c ] && [ \"$2\" = /imexd/imexd.cfg ] && [ \"$3\" = -q ] || exit 64\nprintf '{text}'\nexit {code}\n")
    p.chmod(0o755)
    return str(p)


@pytest.fixture(autouse=True)
def _built(pkg):
    assert os.path.exists(DAEMON)


def test_check_noop_without_clique(tmp_path):
    r = daemon(["check"], {"CLIQUE_ID": "", "FABRIC_PROBE_VERDICT_PATH": str(tmp_path / "none.json")})
    assert r.returncode == 0
    assert r.stdout == "check succeeded (noop, clique ID is empty)\n"  # main.go:437, byte for byte


def test_check_ready_with_clique(tmp_path):
    r = daemon(["check"], {"CLIQUE_ID": "u.1", "CDPROBE_IMEX_CTL": fake_ctl(tmp_path, "READY\\n"),
                           "FABRIC_PROBE_VERDICT_PATH": str(tmp_path / "none.json")})
    assert r.returncode == 0 and r.stdout == ""


@pytest.mark.parametrize("text,code,needle", [
    ("NOT_READY\\n", 0, "IMEX daemon not ready: NOT_READY"),
    ("READY", 0, "IMEX daemon not ready: READY"),               # missing newline is not READY\n
    ("READY\\nmore\\n", 0, "IMEX daemon not ready"),
    ("READY\\n", 3, "IMEX daemon check failed: error running"),  # non-zero exit fails even with READY
])
def test_check_not_ready_with_clique(tmp_path, text, code, needle):
    r = daemon(["check"], {"CLIQUE_ID": "u.1", "CDPROBE_IMEX_CTL": fake_ctl(tmp_path, text, code),
                           "FABRIC_PROBE_VERDICT_PATH": str(tmp_path / "none.json")})
    assert r.returncode == 1 and needle in r.stderr


def test_check_missing_imex_ctl_fails(tmp_path):
    r = daemon(["check"], {"CLIQUE_ID": "u.1", "CDPROBE_IMEX_CTL": str(tmp_path / "absent"),
                           "FABRIC_PROBE_VERDICT_PATH": str(tmp_path / "none.json")})
    assert r.returncode == 1 and "IMEX daemon check failed" in r.stderr


def verdict(tmp_path, ok, **kw):
    d = {"time_unix": kw.pop("time_unix", 2000000000), "ok": ok, "n": 8, "unreachable_pairs": kw.pop("unreachable", 0),
         "min_gbps_read": 665.0, "min_gbps_write": 690.0, "probe_ms": 3.5, "bytes_per_pair": 153391616, "error": ""}
    d.update(kw)
    p = tmp_path / "fabricprobe.json"
    p.write_text(json.dumps(d, indent=1).replace(": True", ": true"))
    return str(p)


def test_check_consults_probe_verdict(tmp_path):
    env = {"CLIQUE_ID": ""}
    assert daemon(["check"], {**env, "FABRIC_PROBE_VERDICT_PATH": verdict(tmp_path, True)}).returncode == 0
    r = daemon(["check"], {**env, "FABRIC_PROBE_VERDICT_PATH": verdict(tmp_path, False, unreachable=2)})
    assert r.returncode == 1 and "fabric probe failed: 2 unreachable pair(s)" in r.stderr
    # the reference's no-op text is still printed first: the IMEX part of the gate did pass
    assert r.stdout == "check succeeded (noop, clique ID is empty)\n"
    # a stale verdict gates only when a maximum age is configured
    old = verdict(tmp_path, True, time_unix=1000)
    assert daemon(["check"], {**env, "FABRIC_PROBE_VERDICT_PATH": old}).returncode == 0
    r = daemon(["check"], {**env, "FABRIC_PROBE_VERDICT_PATH": old, "FABRIC_PROBE_MAX_AGE_S": "60"})
    assert r.returncode == 1 and "stale" in r.stderr


def test_run_requires_cdi_env():
    r = daemon(["run", "--once"], {})
    assert r.returncode == 1
    assert "CDI container edits did not apply -- is CDI enabled in your container runtime?" in r.stderr  # main.go:218


def test_usage():
    assert daemon([]).returncode == 2


@pytest.mark.skipif(gpu_count() > 0, reason="CPU-only behaviour")
def test_run_without_gpu_does_not_gate(tmp_path):
    """No driver: the probe is 'not supported', no verdict is written and check() keeps passing —
    there is no CPU stand-in that would write a fake verdict."""
    v = tmp_path / "fabricprobe.json"
    r = daemon(["run", "--once"], {"COMPUTE_DOMAIN_UUID": "cd-1", "CDPROBE_LIBRARY": LIB,
                                   "FABRIC_PROBE_VERDICT_PATH": str(v)})
    assert r.returncode == 0 and "fabric probe not supported on this node" in r.stderr
    assert not v.exists()
    assert daemon(["check"], {"CLIQUE_ID": "", "FABRIC_PROBE_VERDICT_PATH": str(v)}).returncode == 0


@pytest.mark.gpu
def test_run_once_writes_a_passing_verdict_and_check_reads_it(tmp_path):
    v = tmp_path / "fabricprobe.json"
    # defaults on purpose (1 GiB per GPU, library gate): this is what the daemon pod would run
    m = tmp_path / "fabricprobe.prom"
    env = {"COMPUTE_DOMAIN_UUID": "cd-1", "CDPROBE_LIBRARY": LIB, "FABRIC_PROBE_VERDICT_PATH": str(v),
           "FABRIC_PROBE_METRICS_PATH": str(m)}
    r = daemon(["run", "--once"], env, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "t_fabric_probe" in r.stderr and "fabric probe: verdict ok" in r.stderr
    d = json.loads(v.read_text())
    assert d["ok"] is True and d["unreachable_pairs"] == 0 and d["n"] == gpu_count() and d["probe_ms"] > 0
    n = d["n"]
    assert len(d["reach_read"]) == n * n and all(x == 1 for x in d["reach_read"]) and len(d["gbps_write"]) == n * n
    prom = m.read_text()
    assert "nvidia_dra_fabric_probe_duration_seconds" in prom and "nvidia_dra_fabric_probe_unreachable_pairs 0" in prom
    assert prom.count("nvidia_dra_fabric_probe_pair_gbps{") == 2 * (n * (n - 1) if n > 1 else 1)
    assert daemon(["check"], {"CLIQUE_ID": "", "FABRIC_PROBE_VERDICT_PATH": str(v)}).returncode == 0
