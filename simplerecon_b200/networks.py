"""Parameter container of the matching MLP.

Same constructor, attribute (``.net``) and ``state_dict`` keys
(``net.{0,2,4}.{weight,bias}``) as the reference's ``MLP``
(reference modules/networks.py:129-147) so checkpoints load strictly.  On the hot
path its weights are read by the CUDA kernels directly; ``forward`` is the plain
module semantics for anyone who calls it on its own.
"""
from torch import nn


class MLP(nn.Module):
    def __init__(self, channel_list, disable_final_activation=False):
        super().__init__()
        layers = []
        for c_in, c_out in zip(channel_list[:-1], channel_list[1:]):
            layers.append(nn.Linear(c_in, c_out))
            layers.append(nn.LeakyReLU(inplace=True))
        if disable_final_activation:
            layers = layers[:-1]
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x)
