"""Helpers mirroring the reference's utils.py entries that the model constructors rely on."""
import torch.nn as nn


def ixvr(m):
    """Xavier-normal weights and bias 0.01 for GRU / LSTM / Linear layers (reference utils.py:100-118);
    embeddings and norm layers are left untouched."""
    if isinstance(m, (nn.GRU, nn.LSTM)):
        for name, prm in m.named_parameters():
            if 'weight' in name:
                nn.init.xavier_normal_(prm)
            elif 'bias' in name:
                nn.init.constant_(prm, 0.01)
    elif isinstance(m, nn.Linear):
        nn.init.xavier_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0.01)
