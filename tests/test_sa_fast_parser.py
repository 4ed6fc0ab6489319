"""Host build of kernel 3b's window-based SA piece parser (csrc/exlr_sa_parse.cuh), fuzzed against a plain restatement of
parse_supplementary_alignment / parse_cigar / find_first_match_pos (reference src/utils.rs:12-42, 88-139).
The GPU parity tests exercise the same code on the device; this one runs on the CPU suite and covers far more byte patterns."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("sa_fast") / "sa_fast_harness")
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", out, os.path.join(ROOT, "tests", "sa_fast_harness.cpp")], check=True)
    return out


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_fast_parser_agrees_with_reference_restatement(harness, seed):
    r = subprocess.run([harness, str(seed), "400000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok:")
