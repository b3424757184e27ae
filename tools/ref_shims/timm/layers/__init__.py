import torch

trunc_normal_ = torch.nn.init.trunc_normal_


def get_norm_layer(name):
    return torch.nn.LayerNorm
