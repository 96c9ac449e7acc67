"""bench.py's GPU arm end to end on CPU with stand-ins for the CUDA pieces (tests/bench_dryrun.py): the control flow and
the arithmetic of the JSON line, for both step definitions.  The numbers mean nothing; the contract keys must be there."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("extra,passes", [(["--steps", "4", "--warmup", "2"], 1), (["--steps", "2", "--warmup", "1", "--step", "pass"], 2),
                                          (["--steps", "20", "--warmup", "5", "--verify", "50"], 1)])
def test_bench_main_dry_run(extra, passes):
    r = subprocess.run([sys.executable, os.path.join(HERE, "bench_dryrun.py"), "--config", "dry_24v", "--cpu-sample", "16", "--wave", "16", "--fb-wave", "32"] + extra,
                       capture_output=True, text=True, timeout=900)
    # --verify compares the stand-in engine's verdicts with the oracle: they cannot agree, bench.py must exit with 3
    assert r.returncode == (3 if "--verify" in extra else 0), r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    K = int(extra[1])
    assert d["steps"] == K and d["warmup"] == int(extra[3]) and d["n_gpus"] == 1 and d["edges"] == 276
    assert d["config"]["workload"] == "dry_24v" and "l2" in d["config"] and d["unit"] == "pairs/s"
    # value = pairs of the timed passes / their time
    assert abs(d["value"] - 276 * passes / (d["ms_per_step"] * K * 1e-3)) < 1e-6 * d["value"]
    for key in ("e2e", "roofline", "cpu_baseline", "gpu_launches", "clocks", "host_counters", "branch_mix", "run_config"):
        assert key in d
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] <= d["value"] * 1.0001
    assert d["roofline"]["kernel"] == d["roofline"]["dominant"]
    assert d["cpu_baseline"]["single_thread"]["value"] > 0
    assert len(d["step_wall_ms"]["e2e"]) == K
    if "--verify" in extra:
        assert d["verify"]["tuples"]["tuples"] == 50  # (stand-in verdicts cannot agree with the oracle; the plumbing ran)


def test_bench_main_dry_run_two_ranks():
    """The same under torchrun with two ranks (gloo): shard pinning, batches bracketed by barriers, the record exchange and
    the prefetch gather of builder.py, max-over-ranks timing, one JSON line from rank 0."""
    port = 34500 + (os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(HERE, "bench_dryrun.py"), "--gpus", "2", "--config", "dry_24v",
                        "--cpu-sample", "16", "--wave", "16", "--fb-wave", "32", "--steps", "5", "--warmup", "2"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["steps"] == 5 and d["edges"] == 276
    assert d["host_s_per_pass"]["exchanges"] > 0
    assert abs(d["value"] - 276 / (d["ms_per_step"] * 5 * 1e-3)) < 1e-6 * d["value"]
