"""``ME.utils`` names the reference touches (``networks.py:46``: ``ME.utils.kaiming_normal_``)."""
import torch.nn as nn


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """Kaiming-normal initialisation of a sparse-convolution kernel ``[K^3, Cin, Cout]``."""
    return nn.init.kaiming_normal_(tensor, a=a, mode=mode, nonlinearity=nonlinearity)
