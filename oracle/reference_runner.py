"""Runs the UNMODIFIED reference modules (bytecode in oracle/_ref, built by oracle/build_ref.py) on seeded
synthetic workloads.  TEST / BASELINE INFRASTRUCTURE ONLY: imported by tests/, bench.py's reference / cpu_baseline
legs and oracle/gen_golden_full.py — never by the product package.

What is executed is the reference's own code: `model.S2VTAttModel.S2VTAttModel(...)(vid_feats, s)`,
`train_utils.calc_masked_loss`, `calc_masked_accuracy`, `loss.backward()` — i.e. `run_iter` of train.py:32-44 plus
train.py:157-158 (train.py itself needs tensorboardX / nlgeval, absent from the image, so its 10 lines are restated
in `run_iter` below).
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "model", "S2VTAttModel.bc"))


class _RefFinder(importlib.abc.MetaPathFinder):
    """Imports `utils`, `train_utils`, `model` and `model.*` from the bytecode files oracle/build_ref.py wrote."""
    NAMES = ("utils", "train_utils", "model")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] not in self.NAMES:
            return None
        rel = fullname.replace(".", os.sep)
        pkg = os.path.join(REF, rel, "__init__.bc")
        mod = os.path.join(REF, rel + ".bc")
        if os.path.exists(pkg):
            loader = importlib.machinery.SourcelessFileLoader(fullname, pkg)
            return importlib.util.spec_from_file_location(fullname, pkg, loader=loader,
                                                          submodule_search_locations=[os.path.join(REF, rel)])
        if os.path.exists(mod):
            loader = importlib.machinery.SourcelessFileLoader(fullname, mod)
            return importlib.util.spec_from_file_location(fullname, mod, loader=loader)
        return None


_mods = None


def modules():
    """-> dict of the reference modules (utils, train_utils, model.*) imported from oracle/_ref bytecode."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    clash = [m for m in _RefFinder.NAMES if m in sys.modules
             and not str(getattr(sys.modules[m], "__file__", "")).startswith(REF)]
    if clash:
        raise RuntimeError("modules %s already imported from elsewhere" % clash)
    finder = _RefFinder()
    sys.meta_path.insert(0, finder)
    try:
        names = ["utils", "train_utils", "model.S2VTModel", "model.S2VTAttModel", "model.RationaleNet",
                 "model.SpatialNet"]
        _mods = {n: importlib.import_module(n) for n in names}
    finally:
        sys.meta_path.remove(finder)
    return _mods


class FakeGlove:
    """Duck type of utils.GloveLoader (utils.py:52-98): the four special tokens are the last four ids."""

    def __init__(self, vocab, embed, vectors=None):
        self.vocab = vocab
        self.word_vectors = [np.zeros(embed, np.float32) for _ in range(vocab)] if vectors is None else list(vectors)

    def get_id(self, w):
        return {"<sos>": self.vocab - 4, "<eos>": self.vocab - 3, "<pad>": self.vocab - 2, "<unk>": self.vocab - 1}[w]


def set_device(device):
    """train_utils.py:8-9 binds a module-level `device` at import ('cuda' when one is visible); the CPU arm on a GPU
    box needs it to say 'cpu' (equivalent to running the reference with CUDA_VISIBLE_DEVICES='')."""
    import torch
    modules()["train_utils"].device = torch.device(device)


def build_s2vtatt(dims, params=None, dropout_p=0.0, device="cpu", dtype=None):
    """Reference S2VTAttModel with (optionally) a given state_dict (numpy arrays keyed like the state_dict)."""
    import torch
    B, N, V, H, E, L, Vc = dims
    m = modules()["model.S2VTAttModel"].S2VTAttModel(FakeGlove(Vc, E), dropout_p, H, V, L)
    if params is not None:
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    if dtype is not None:
        m = m.to(dtype)
    return m.to(device)


def run_iter(model, vid, s, s_len, backward=True):
    """train.py:32-44 (+ :157-158): logits = model(vid, s); pred; masked loss; masked accuracy; loss.backward()."""
    import torch
    tu = modules()["train_utils"]
    criterion = torch.nn.CrossEntropyLoss(reduction="none")
    logits = model(vid, s)
    if isinstance(logits, tuple):
        logits = logits[0]
    pred = torch.argmax(logits, dim=2)
    loss = calc = tu.calc_masked_loss(logits, s, s_len, criterion)
    acc = tu.calc_masked_accuracy(logits, s, s_len)
    if backward:
        model.zero_grad()
        calc.backward()
    return loss, acc, pred, logits
