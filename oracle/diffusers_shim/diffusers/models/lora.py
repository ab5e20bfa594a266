from torch import nn


class LoRACompatibleLinear(nn.Linear):
    pass
