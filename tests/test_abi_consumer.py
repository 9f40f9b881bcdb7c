"""The C ABI used from plain C: tests/abi_consumer.c (CUDA runtime + include/dlmcq.h only - no Python, no torch in
that process) drives libdlmcq.so and checks it against the plain-C oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "abi_consumer")


def test_consumer_source_uses_only_the_c_abi():
    src = open(os.path.join(ROOT, "tests", "abi_consumer.c")).read()
    assert '#include "../include/dlmcq.h"' in src and "torch" not in src.replace("no torch", "") and "Python.h" not in src


@pytest.mark.gpu
def test_plain_c_host_matches_the_c_oracle():
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build_abi_consumer()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout and r.stdout.count("PASS ") >= 16, r.stdout
