"""CPU: the C-ABI library loads and exports every symbol include/pvcr_b200.h declares, and the ctypes signature table
matches the header's argument counts (no compute calls: there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pvcr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|void|const char\*)\s+(pvcr_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_header_declares_the_expected_surface():
    fns = _header_functions()
    for name in ("pvcr_s2vtatt_fwd", "pvcr_s2vtatt_bwd", "pvcr_s2vt_fwd", "pvcr_s2vt_bwd", "pvcr_generator_fwd",
                 "pvcr_generator_bwd", "pvcr_vocab_ce_fwd", "pvcr_vocab_ce_bwd", "pvcr_s2vtatt_greedy",
                 "pvcr_s2vt_decode_steps", "pvcr_masked_ce", "pvcr_linear_fwd", "pvcr_linear_bwd"):
        assert name in fns, name


def test_library_exports_every_declared_symbol():
    import pvcr_b200  # noqa: F401
    from pvcr_b200 import _lib
    L = _lib.lib()
    for name in _header_functions():
        assert hasattr(L, name), "libpvcr_b200.so does not export %s" % name
    assert L.pvcr_version() >= 100
    assert L.pvcr_prof_num_classes() > 0


def test_ctypes_table_matches_header():
    from pvcr_b200 import _lib
    fns = _header_functions()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert name in fns, "%s bound in _lib.py but not declared in the header" % name
        assert len(argtypes) == fns[name], (name, len(argtypes), fns[name])
    unbound = set(fns) - set(_lib.SIGNATURES) - {"pvcr_last_error", "pvcr_version"}
    assert not unbound, unbound


def test_struct_layouts_match_header():
    """Field order of the ctypes structs == field order of the C typedefs."""
    from pvcr_b200 import _lib
    src = open(os.path.join(ROOT, "include", "pvcr_b200.h")).read()
    for cname, fields in (("PvcrS2vtAttParams", _lib.ATT_PARAM_FIELDS), ("PvcrS2vtParams", _lib.S2VT_PARAM_FIELDS),
                          ("PvcrGenParams", _lib.GEN_PARAM_FIELDS)):
        end = src.index("} %s;" % cname)
        body = src[src.rindex("typedef struct {", 0, end):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = re.findall(r"const float\*\s*(\w+)\s*;", body)
        assert names == list(fields), (cname, names)


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pvcr_b200 import functional as F_
    with pytest.raises(AssertionError, match="no CPU path"):
        F_._f32c(torch.zeros(2))


def test_spatial_encode_host_side():
    """pvcr_spatial_encode_*: the workspace query is host arithmetic (grows with the frame count, covers the saved activations), and
    a null argument is refused with an error code and message before anything touches the device."""
    import pvcr_b200  # noqa: F401
    from pvcr_b200 import _lib
    L = _lib.lib()
    B, Kc, H, F = 128, 36, 512, 2048
    w40, w41 = L.pvcr_spatial_encode_workspace(B, 40, Kc, H, F, 1), L.pvcr_spatial_encode_workspace(B, 41, Kc, H, F, 1)
    per_frame = B * (4 * H + 4 * H + F + 3 * H + 4 * H + H) * 4       # qgh, saved gates, ctx | dgi, d1, dv partials (fp32)
    assert w41 - w40 >= per_frame and w41 - w40 < 2 * per_frame + (1 << 20)
    assert L.pvcr_spatial_encode_workspace(B, 40, Kc, H, F, 3) > w40          # split planes are wider
    rc = L.pvcr_spatial_encode_fwd(B, 40, Kc, H, F, 1, None, None, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc != 0 and b"null argument" in L.pvcr_last_error()
    rc = L.pvcr_spatial_encode_bwd(B, 40, Kc, H, F, 1, *([None] * 16), None, 0, None)
    assert rc != 0 and b"null argument" in L.pvcr_last_error()
