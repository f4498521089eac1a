"""Helpers for the GPU parity tests: build the drop-in modules from a golden fixture's weights."""
import numpy as np
import torch

from tests.golden_util import load


class FixtureGlove:
    """Duck type of the reference GloveLoader (utils.py:52-66): .word_vectors and .get_id()."""

    def __init__(self, vocab, embed):
        self.word_vectors = [np.zeros(embed, np.float32) for _ in range(vocab)]
        self.vocab = vocab

    def get_id(self, w):
        return {"<sos>": self.vocab - 4, "<eos>": self.vocab - 3, "<pad>": self.vocab - 2, "<unk>": self.vocab - 1}[w]


def load_case(tag):
    d, params, grads = load(tag)
    B, N, V, H, E, L, Vc = (int(x) for x in d["dims"])
    return d, params, grads, (B, N, V, H, E, L, Vc)


def to_cuda(model, params):
    sd = {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in params.items()}
    model.load_state_dict(sd)
    return model.cuda()


def grads_of(model):
    return {k: p.grad.detach().double().cpu().numpy() for k, p in model.named_parameters() if p.grad is not None}
