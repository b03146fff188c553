"""Minimal ``MetaTemplate`` (reference methods/meta_template.py:10-18): the six attributes DKT relies on.
The generic train/test loops of the other few-shot methods are outside the DKT hot path."""
import torch.nn as nn


class MetaTemplate(nn.Module):
    def __init__(self, model_func, n_way, n_support, change_way=True):
        super().__init__()
        self.n_way = n_way
        self.n_support = n_support
        self.n_query = -1          # changes with the input
        self.feature = model_func()
        self.feat_dim = self.feature.final_feat_dim
        self.change_way = change_way

    def set_forward(self, x, is_feature=False):
        pass

    def set_forward_loss(self, x):
        pass

    def forward(self, x):
        return self.feature.forward(x)
