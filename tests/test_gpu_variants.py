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


def _run_decode(env):
    sel = [os.path.join(HERE, "test_gpu_fullsize_reference.py") + "::test_full_size_greedy_ids_vs_reference",
           os.path.join(HERE, "test_gpu_decode.py") + "::test_greedy_batch_and_width_sweep_vs_oracle",
           os.path.join(HERE, "test_gpu_decode.py") + "::test_s2vtatt_greedy_ids_bit_exact"]
    r = subprocess.run([sys.executable, "-m", "pytest"] + sel + ["-m", "gpu", "-x", "-q"], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=1200, cwd=os.path.dirname(HERE))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert " passed" in r.stdout


@pytest.mark.parametrize("env", [
    # the decode loop as it was before the split3 kernel: generic six-plane GEMMs, split-K [q | gh] product, context GEMM
    # per step, separate arg-max combine launch, no programmatic launches
    {"PVCR_ARGMAX_SPLIT3": "0", "PVCR_NO_SPLIT3_STORE": "1", "PVCR_NO_DECODE_FOLD_ARGMAX": "1", "PVCR_NO_DECODE_PC": "1",
     "PVCR_NO_DECODE_PDL": "1"},
    # split3 kernel on 32-column K chunks (64-byte swizzle rows) for both the vocabulary and the [q | gh] product
    {"PVCR_ARGMAX_BK": "32", "PVCR_SPLIT3_STORE_BK": "32"},
    # 160-column vocabulary tiles on 144 CTAs, W_v planes not K-blocked, L2 prefetch ahead of the ring
    {"PVCR_ARGMAX_SPLIT3": "160", "PVCR_ARGMAX_CTAS": "144", "PVCR_NO_WV_BLOCKED": "1", "PVCR_WV_PREFETCH": "1"},
], ids=["pre_split3_loop", "split3_bk32", "split3_bn160_unblocked_prefetch"])
def test_decode_variant_ids_bit_exact(env):
    """Greedy ids stay bit-exact against the reference fixture (full cfg5 size) and the oracle under every decode-loop knob."""
    _run_decode(env)


def test_conv_taps_one_launch_per_tap_variant():
    """The nine-launch convolution (what the split-precision modes and odd channel counts use) under the single-plane mode."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_spatial_front.py"), "-m", "gpu", "-x", "-q", "-k",
                        "conv_bn_relu_front"], env=dict(os.environ, PVCR_NO_CONV_FUSED_TAPS="1"), capture_output=True, text=True,
                       timeout=1200, cwd=os.path.dirname(HERE))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
