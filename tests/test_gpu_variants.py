"""Kernel variants that are off by default (measured slower, kept in the tree with their A/B knobs) stay parity-green:
the full-size reference comparison is re-run in a subprocess with the knob set (knobs are read once per process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(env, select):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_fullsize_reference.py") + "::" + select,
                        "-m", "gpu", "-x", "-q"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=1200,
                       cwd=os.path.dirname(HERE))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert " passed" in r.stdout


@pytest.mark.parametrize("env", [
    {"PVCR_GRU_CLUSTER": "1"},                                   # encoder forward sweep with the DSMEM (cluster) exchange
    {"PVCR_DEC_FWD_XCHG": "0", "PVCR_DEC_BWD_XCHG": "2"},        # q by group barrier, dctx polled by warp 0
    {"PVCR_NO_TMA_XCHG": "1"},                                   # cp.async exchange loads in every sweep
], ids=["gru_cluster", "exchange_modes", "no_tma"])
def test_variant_matches_reference_fixture(env):
    _run(env, "test_full_cfg2_vs_reference_fixture[bf16]")
