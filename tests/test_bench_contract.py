"""CPU: the arithmetic behind bench.py's roofline line (algorithmic bytes per launch) against the oracle's block geometry and
DESIGN.md's stated figures, and the command-line contract the driver relies on.  No GPU, no timing."""
import importlib
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import oracle as O

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def bench():
    sys.path.insert(0, str(ROOT))
    return importlib.import_module("bench")


def test_block_geometry_matches_the_oracle(bench, oracle):
    for t, (qk, ts) in bench.BLOCK.items():
        assert oracle.blck_size(t) == qk, bench.TNAME[t]
        for K in (256, 4096, 14336):
            assert oracle.row_size(t, K) == bench.row_bytes(t, K) == K // qk * ts


def test_algorithmic_bytes_are_the_documented_figures(bench):
    # Q4_K 4096 -> 11008, n = 1: weights + x + y
    assert bench.weight_bytes(4096, 11008) == 11008 * 16 * 144 == 25_362_432
    assert bench.algorithmic_bytes(4096, 11008, 1) == 25_422_848
    assert bench.algorithmic_bytes(4096, 4096, 1, t=2) == 9_469_952
    # the headline workload is BASELINE.json configs[1] and the sweep is larger than two L2s
    assert (bench.WL["K"], bench.WL["M"], bench.WL["N"], bench.WL["type_id"]) == (4096, 11008, 1, O.Q4_K)
    assert bench.NBUF * bench.weight_bytes(4096, 11008) > 2 * 126e6
    assert bench.SWEEPS_PER_STEP % bench.SWEEPS_PER_GRAPH == 0


def test_measured_peaks_or_the_profiling_guide_fallback(bench, tmp_path, monkeypatch):
    p = bench.measured_peaks()
    assert p["hbm"] > 1000 and p["tc"] > 100 and p["tc_burst"] >= p["tc"] and isinstance(p["src"], str)
    monkeypatch.setattr(bench, "ROOT", tmp_path)                    # no MEASURED_PEAKS.json: the guide's stated fallback, and it says so
    f = bench.measured_peaks()
    assert f["src"].startswith("fallback") and f["hbm"] == 6650.0


def test_command_line_contract():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
