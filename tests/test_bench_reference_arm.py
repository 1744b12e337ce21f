"""bench.py --impl reference: runs without a GPU, prints the contract's JSON line, and keeps the CUDA library out of
its process (the arm asserts that on /proc/self/maps itself; here the line and the exit code are checked)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-spp", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip().splitlines()


def test_reference_arm_line():
    lines = _run({"OMP_NUM_THREADS": "1"})  # what torchrun exports: the arm must still use every host thread
    assert len(lines) == 1  # stdout is the JSON line and nothing else
    line = json.loads(lines[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mpaths/s" and line["unit"] == "Mpaths/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["config"]["scene"] == "cornell" and line["config"]["width"] == 600 and line["config"]["spp"] == 1000
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    want = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == want, (line["cpu_baseline"]["cores"], want)
    assert line["e2e"] == {"value": line["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_stdout_carries_only_the_json_line():
    """What a library prints to file descriptor 1 during the run (NCCL's version banner under torchrun at 8 GPUs did)
    must not land in front of the JSON: bench.claim_stdout points fd 1 at stderr and emit() writes to the descriptor
    stdout had."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('a stray print'); bench.emit({'metric': 'x'})" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout == '{"metric": "x"}\n'
    assert "NCCL version" in p.stderr and "a stray print" in p.stderr
