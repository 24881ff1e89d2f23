"""Stand-in for ``espnet2.torch_utils.get_layer_from_string`` (espnet==202308, requirements2.txt:17).

TEST INFRASTRUCTURE ONLY.  The reference only ever asks for ``get_layer("prelu")`` (tfgridnet_causal.py:646).
"""
import torch.nn as nn


def get_layer(name):
    table = {"prelu": nn.PReLU, "relu": nn.ReLU, "elu": nn.ELU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
    return table[name.lower()]
