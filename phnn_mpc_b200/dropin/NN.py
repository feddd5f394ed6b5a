"""MLP container with the reference's layout (src/NN.py:6-40): ``self.net`` is a Sequential of
Linear [LayerNorm] activation [Dropout] ... Linear, so checkpoints keyed ``*.net.<i>.weight``
load unchanged, and the parameter initialisation consumes the RNG in the same order."""
import math

import torch.nn as nn


class MLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_sizes=(128, 128), activation=nn.SiLU, dropout=0.0,
                 layer_norm=False, bias=True):
        super().__init__()
        self.activation_name = activation.__name__
        self.uses_layer_norm = bool(layer_norm)
        self.dropout_p = float(dropout)
        widths = [input_dim] + list(hidden_sizes)
        mods = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            mods.append(nn.Linear(fan_in, fan_out, bias=bias))
            if layer_norm:
                mods.append(nn.LayerNorm(fan_out))
            mods.append(activation())
            if dropout > 0:
                mods.append(nn.Dropout(dropout))
        mods.append(nn.Linear(widths[-1], output_dim, bias=bias))
        self.net = nn.Sequential(*mods)
        # second pass over the Linear layers, as the reference does after construction
        for lin in (m for m in self.net if isinstance(m, nn.Linear)):
            nn.init.kaiming_uniform_(lin.weight, a=math.sqrt(5))
            if lin.bias is not None:
                fan_in = lin.weight.shape[1]
                bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
                nn.init.uniform_(lin.bias, -bound, bound)

    def kernel_compatible(self):
        """True when the CUDA kernels implement this stack: Tanh, no LayerNorm.  Dropout layers are accepted: they are the
        identity in eval mode, which is the mode the controllers put the model in (src/mpc_controller.py:44,
        src/mpc_controller_canonical.py:54); ``check_mode`` refuses a forward in training mode.  ``bias=False`` is a
        zero bias for the kernels."""
        return self.activation_name == "Tanh" and not self.uses_layer_norm

    def check_mode(self):
        if self.dropout_p > 0.0 and self.training:
            raise RuntimeError("this MLP has Dropout(p=%g): the CUDA path evaluates the eval-mode function only -- call "
                               "model.eval() first (the reference's controllers do, src/mpc_controller.py:44)" % self.dropout_p)

    def forward(self, x):
        return self.net(x)
