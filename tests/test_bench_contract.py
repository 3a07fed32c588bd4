"""bench.py contract checks that need no GPU: the reference arm's JSON line (bounded CPU sample of
the device arm's workload) and its behaviour under a torchrun-style multi-process launch."""
import json
import os
import subprocess
import sys

import pytest

from parity_util import ROOT

BENCH = os.path.join(ROOT, "bench.py")
REF_EXE = os.path.join(ROOT, "meep_b200", "lib", "bench_ref_f64")
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _run(extra, env=None):
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "2", "--warmup", "3",
                        "--cpu-n", "32"] + extra, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="reference bench driver not built")
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run([]).strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "Yee cell-updates/s" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same workload text as the device arm prints for the default configuration
    # (one GPU runs the 1024^3 cell of the scaling series; configs[1] = 512^3 rides along in the device arm)
    assert "1024x1024x1024" in d["config"]["workload"] and d["scaling"] == "strong"


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="reference bench driver not built")
def test_reference_arm_under_a_two_rank_launch_only_rank_zero_reports():
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29713")
        outs.append(_run(["--gpus", "2"], env=env).strip())
    assert outs[1] == ""
    d = json.loads(outs[0])
    # N > 1: the north-star case (BASELINE.json configs[4]): 1024^3, strong scaling
    assert d["n_gpus"] == 2 and "1024x1024x1024" in d["config"]["workload"] and d["scaling"] == "strong"


def test_max_over_ranks_reduction_with_gloo_world_size_2(tmp_path):
    """the timing reduction bench.py uses (max over ranks), exercised over gloo on the CPU"""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys\n"
        "sys.path.insert(0, %r)\n"
        "import torch.distributed as dist\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank()\n"
        "v = bench.dist_max(10.0 + 5.0 * r, dist.get_world_size(), 'cpu')\n"
        "dist.barrier()\n"
        "assert v == 15.0, v\n"
        "print('ok', r)\n" % ROOT)
    procs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29714")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0 and "ok" in out, out[-2000:]
