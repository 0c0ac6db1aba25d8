"""N > 1 host path on CPU: world_size-2 gloo rank group (what bench.py uses to bracket timed
regions and take max-over-ranks) + the reference arm's rank discipline under torchrun."""
import json
import os
import socket
import subprocess
import sys
import textwrap

from conftest import ROOT


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CHILD = textwrap.dedent(
    """
    import os, sys, json
    sys.path.insert(0, %r)
    import cdprobe_pkg
    m = cdprobe_pkg.load()
    g = m.distutil.RankGroup(backend="gloo")
    g.barrier()
    lib = m.abi.load_library()
    rc = lib.cdprobe_rendezvous_selftest(g.session("t").encode(), g.rank, g.world, 20000)
    out = {"rank": g.rank, "world": g.world, "max": g.max(10.0 + g.rank), "min": g.min(10.0 + g.rank),
           "session": g.session(), "rdv": rc, "uuids": g.gather_objects(f"GPU-{g.rank:04d}")}
    g.close()
    print("OUT " + json.dumps(out))
    """
) % ROOT


def launch(world, argv, extra_env=None):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), TORCHELASTIC_RUN_ID="cpu-test")
        env.update(extra_env or {})
        procs.append(subprocess.Popen(argv, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    return [p.communicate(timeout=180) + (p.returncode,) for p in procs]


def test_gloo_rank_group_world2(pkg):
    res = launch(2, [sys.executable, "-c", CHILD])
    outs = []
    for so, se, rc in res:
        assert rc == 0, se[-1500:]
        outs.append(json.loads([l for l in so.splitlines() if l.startswith("OUT ")][-1][4:]))
    assert sorted(o["rank"] for o in outs) == [0, 1]
    for o in outs:
        assert o["world"] == 2 and o["max"] == 11.0 and o["min"] == 10.0 and o["rdv"] == 0
        assert o["uuids"] == ["GPU-0000", "GPU-0001"]  # rank order on every rank: bench.py's UUID -> NVML index matching
    assert outs[0]["session"] == outs[1]["session"]  # every rank derives the same rendezvous name


def test_reference_arm_only_rank0_prints(oracle, tmp_path):
    """`bench.py --impl reference` under a 2-rank launch: rank 0 times the NVML poll (here against the
    fake NVML) and prints one JSON line; the other rank exits 0 without work."""
    sc = tmp_path / "sc.txt"
    sc.write_text("gpus 2\n")
    fake = os.path.join(ROOT, "tests", "fake_nvml", "libnvidia-ml.so.1")
    res = launch(2, [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                     "3", "--warmup", "1"], {"CDORACLE_NVML_PATH": fake, "FAKE_NVML_SCENARIO": str(sc)})
    lines = []
    for so, se, rc in res:
        assert rc == 0, se[-1500:]
        lines += [l for l in so.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "nvlink_probe_ms" and j["unit"] == "ms"
    assert j["higher_is_better"] is False and j["n_gpus"] == 2 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] in (1, 2)
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["reach_all_ones"] is True
