"""ctypes loader for libpvcr_b200.so (the C ABI in include/pvcr_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PVCR_LIB", os.path.join(_HERE, "libpvcr_b200.so"))

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_size = ctypes.c_size_t
c_vp = ctypes.c_void_p
c_f = ctypes.c_float
c_u64 = ctypes.c_uint64

_lib = None


class PvcrError(RuntimeError):
    pass


class PvcrDims(ctypes.Structure):
    _fields_ = [("B", c_int), ("N", c_int), ("V", c_int), ("H", c_int), ("E", c_int), ("L", c_int), ("Vc", c_int),
                ("nsplit", c_int), ("dropout_p", c_f), ("seed", c_u64)]


ATT_PARAM_FIELDS = ["enc_w_ih", "enc_w_hh", "enc_b_ih", "enc_b_hh", "emb", "dec_w_ih", "dec_w_hh", "dec_b_ih",
                    "dec_b_hh", "att_wk", "att_wq", "att_v", "out_w", "out_b"]
S2VT_PARAM_FIELDS = ["emb", "rnn1_w_ih", "rnn1_w_hh", "rnn1_b_ih", "rnn1_b_hh", "rnn2_w_ih", "rnn2_w_hh", "rnn2_b_ih",
                     "rnn2_b_hh", "out_w", "out_b"]
FRONT_PARAM_FIELDS = ["conv1_w", "conv1_b", "bn1_w", "bn1_b", "conv2_w", "conv2_b", "bn2_w", "bn2_b"]
GEN_PARAM_FIELDS = ["w_ih", "w_hh", "b_ih", "b_hh", "w_ih_r", "w_hh_r", "b_ih_r", "b_hh_r", "lin_w", "lin_b"]


def _ptr_struct(name, fields):
    return type(name, (ctypes.Structure,), {"_fields_": [(f, c_vp) for f in fields]})


PvcrS2vtAttParams = _ptr_struct("PvcrS2vtAttParams", ATT_PARAM_FIELDS)
PvcrS2vtAttGrads = _ptr_struct("PvcrS2vtAttGrads", ATT_PARAM_FIELDS)
PvcrS2vtParams = _ptr_struct("PvcrS2vtParams", S2VT_PARAM_FIELDS)
PvcrS2vtGrads = _ptr_struct("PvcrS2vtGrads", S2VT_PARAM_FIELDS)
PvcrGenParams = _ptr_struct("PvcrGenParams", GEN_PARAM_FIELDS)
PvcrGenGrads = _ptr_struct("PvcrGenGrads", GEN_PARAM_FIELDS)
PvcrSpatialFrontParams = _ptr_struct("PvcrSpatialFrontParams", FRONT_PARAM_FIELDS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PvcrError("%s not found: build it with `python __graft_entry__.py` (no CPU fallback exists)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pvcr_last_error.restype = ctypes.c_char_p
        _lib.pvcr_version.restype = c_int
        _declare(_lib)
    return _lib


def _sig(L, name, restype, argtypes):
    fn = getattr(L, name)
    fn.restype = restype
    fn.argtypes = argtypes


# name -> (restype, argtypes); mirrors include/pvcr_b200.h one to one (tests/test_abi.py cross-checks the header)
P = ctypes.POINTER
SIGNATURES = {
    "pvcr_set_seed_step": (None, [c_vp]),
    "pvcr_side_mode": (c_int, [c_int]),
    "pvcr_side_join": (c_int, [c_vp]),
    "pvcr_side_join_lane": (c_int, [c_vp, c_int]),
    "pvcr_side_wait_milestone": (c_int, [c_vp, c_int]),
    "pvcr_debug_phase_timing": (c_int, [c_int]),
    "pvcr_debug_phase_read": (c_int, [P(ctypes.c_longlong), c_int]),
    "pvcr_prof_num_classes": (c_int, []),
    "pvcr_prof_class_name": (ctypes.c_char_p, [c_int]),
    "pvcr_prof_enable": (None, [c_int]),
    "pvcr_prof_reset": (None, []),
    "pvcr_prof_read": (c_int, [P(c_u64), P(ctypes.c_double), P(ctypes.c_double)]),
    "pvcr_prof_timeline": (c_int, [P(c_int), P(ctypes.c_float), P(ctypes.c_float), c_int]),
    "pvcr_prof_launch_list": (c_int, [P(c_int), P(ctypes.c_float), P(ctypes.c_double), c_int]),
    "pvcr_linear_fwd_workspace": (c_size, [c_int, c_int, c_int, c_int]),
    "pvcr_linear_fwd": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_size,
                                c_vp]),
    "pvcr_linear_bwd_workspace": (c_size, [c_int, c_int, c_int, c_int]),
    "pvcr_linear_bwd": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_int, c_int,
                                c_int, c_int, c_int, c_vp, c_size, c_vp]),
    "pvcr_wgrad_mn_workspace": (c_size, [c_int, c_int, c_int]),
    "pvcr_wgrad_mn": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_size, c_vp]),
    "pvcr_s2vtatt_workspace": (c_size, [P(PvcrDims), c_int]),
    "pvcr_s2vtatt_fwd": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "pvcr_s2vtatt_bwd": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_vp, c_vp, c_vp, P(PvcrS2vtAttGrads),
                                 c_vp, c_vp, c_size, c_vp]),
    "pvcr_s2vt_workspace": (c_size, [P(PvcrDims), c_int]),
    "pvcr_s2vt_fwd": (c_int, [P(PvcrDims), P(PvcrS2vtParams), c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "pvcr_s2vt_bwd": (c_int, [P(PvcrDims), P(PvcrS2vtParams), c_vp, c_vp, c_vp, c_vp, c_vp, P(PvcrS2vtGrads), c_vp, c_vp,
                              c_size, c_vp]),
    "pvcr_s2vtatt_greedy_workspace": (c_size, [P(PvcrDims)]),
    "pvcr_s2vtatt_greedy": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_size,
                                    c_vp]),
    "pvcr_s2vtatt_beam_workspace": (c_size, [P(PvcrDims), c_int]),
    "pvcr_s2vtatt_beam": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_size,
                                  c_vp]),
    "pvcr_s2vt_decode_steps_workspace": (c_size, [P(PvcrDims)]),
    "pvcr_s2vt_decode_steps": (c_int, [P(PvcrDims), P(PvcrS2vtParams), c_vp, c_vp, c_i64, c_vp, P(ctypes.c_int32), c_f,
                                       c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "pvcr_generator_workspace": (c_size, [P(PvcrDims)]),
    "pvcr_generator_fwd": (c_int, [P(PvcrDims), P(PvcrGenParams), c_vp, c_vp, c_f, c_int, c_vp, c_vp, c_vp, c_vp, c_size,
                                   c_vp]),
    "pvcr_generator_bwd": (c_int, [P(PvcrDims), P(PvcrGenParams), c_vp, c_f, c_vp, c_vp, c_vp, P(PvcrGenGrads), c_vp,
                                   c_size, c_vp]),
    "pvcr_masked_ce": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                               c_vp]),
    "pvcr_rationale_penalties": (c_int, [c_vp, c_int, c_int, c_vp, c_vp]),
    "pvcr_rationale_penalties_bwd": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "pvcr_s2vtatt_bwd_part": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_vp, c_vp, c_vp, P(PvcrS2vtAttGrads),
                                      c_vp, c_vp, c_size, c_vp, c_int]),
    "pvcr_adam_clip_step": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_vp, c_i64, c_vp, c_vp,
                                    c_vp]),
    "pvcr_vocab_ce_workspace": (c_size, [c_int, c_int, c_int, c_int, c_int, c_f]),
    "pvcr_vocab_ce_prepare": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_size, c_vp]),
    "pvcr_vocab_ce_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f, c_u64, c_vp,
                                  c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_size, c_vp]),
    "pvcr_s2vtatt_decode_fwd": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "pvcr_s2vtatt_decode_bwd": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_vp, P(PvcrS2vtAttGrads), c_vp, c_vp,
                                        c_vp, c_size, c_vp]),
    "pvcr_s2vt_decode_fwd": (c_int, [P(PvcrDims), P(PvcrS2vtParams), c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "pvcr_s2vt_decode_bwd": (c_int, [P(PvcrDims), P(PvcrS2vtParams), c_vp, c_vp, c_vp, P(PvcrS2vtGrads), c_vp, c_vp, c_vp,
                                     c_size, c_vp]),
    "pvcr_s2vtatt_greedy_ex": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                       c_size, c_int, c_vp]),
    "pvcr_s2vtatt_decode_greedy": (c_int, [P(PvcrDims), P(PvcrS2vtAttParams), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                           c_size, c_vp]),
    "pvcr_s2vt_decode_greedy": (c_int, [P(PvcrDims), P(PvcrS2vtParams), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_size, c_vp]),
    "pvcr_gru_step_workspace": (c_size, [c_int, c_int, c_int, c_int]),
    "pvcr_gru_step_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_size,
                                  c_vp]),
    "pvcr_gru_step_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                  c_vp, c_vp, c_int, c_vp, c_size, c_vp]),
    "pvcr_spatial_front_workspace": (c_size, [c_int, c_int, c_int, c_int, c_int]),
    "pvcr_spatial_front_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_vp, P(PvcrSpatialFrontParams), c_vp, c_vp, c_vp, c_vp,
                                       c_int, c_f, c_f, c_vp, c_vp, c_vp, c_size, c_vp]),
    "pvcr_spatial_front_bwd": (c_int, [c_int, c_int, c_int, c_int, c_int, P(PvcrSpatialFrontParams), c_int, c_vp,
                                       P(PvcrSpatialFrontParams), c_vp, c_size, c_vp]),
    "pvcr_spatial_attn_fwd": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "pvcr_spatial_attn_bwd": (c_int, [c_int, c_int, c_int, c_int, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp]),
    "pvcr_spatial_encode_workspace": (c_size, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "pvcr_spatial_encode_fwd": (c_int, [c_int] * 6 + [c_vp] * 10 + [c_vp, c_size, c_vp]),
    "pvcr_spatial_encode_bwd": (c_int, [c_int] * 6 + [c_vp] * 16 + [c_vp, c_size, c_vp]),
    "pvcr_out_dropout_apply": (c_int, [c_vp, c_vp, c_i64, c_f, c_u64, c_vp]),
    "pvcr_debug_philox_minmax": (c_int, [c_u64, c_u64, c_u64, c_vp, c_vp]),
    "pvcr_vocab_ce_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f, c_u64, c_vp, c_vp,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
}


def _declare(L):
    for name, (res, args) in SIGNATURES.items():
        _sig(L, name, res, args)


def check(rc, what):
    if rc != 0:
        raise PvcrError("%s failed (%d): %s" % (what, rc, lib().pvcr_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def prof_launch_list(cap=4096):
    """[(class name, ms, work)] for every event-timed launch since the last pvcr_prof_reset(), in enqueue order."""
    L = lib()
    cls = (c_int * cap)()
    ms = (ctypes.c_float * cap)()
    work = (ctypes.c_double * cap)()
    n = L.pvcr_prof_launch_list(cls, ms, work, cap)
    if n < 0:
        check(n, "pvcr_prof_launch_list")
    return [(L.pvcr_prof_class_name(cls[i]).decode(), float(ms[i]), float(work[i])) for i in range(n)]


def prof_read():
    """{class name: (launches, summed event ms, work)} since the last pvcr_prof_reset()."""
    L = lib()
    n = L.pvcr_prof_num_classes()
    launches = (c_u64 * n)()
    ms = (ctypes.c_double * n)()
    work = (ctypes.c_double * n)()
    check(L.pvcr_prof_read(launches, ms, work), "pvcr_prof_read")
    return {L.pvcr_prof_class_name(i).decode(): (int(launches[i]), float(ms[i]), float(work[i])) for i in range(n)}
