"""CPU: unit test of the leader/follower request combiner (redisearch_b200/csrc/micro_batcher.h) that maps concurrent
single-query callers onto the batched entry points.  Compiled with g++ and run as a plain host program."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_micro_batcher_combines_concurrent_callers(tmp_path):
    exe = tmp_path / "mb_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "cpp", "micro_batcher_test.cpp"), "-o", str(exe)],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(": ok") == 3, r.stdout
