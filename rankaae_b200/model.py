"""Model construction surface of the hot path (mirrors `sc/clustering/model.py` of the reference):
`FCEncoder` (model.py:330-378), `FCDecoder` (:518-570), `DiscriminatorFC` (:631-663) and
`GradientReversalLayer` (:8-22), with the same constructor keywords, the same `main`
nn.Sequential structure and therefore the same state_dict keys / parameters() order.

These modules are CONTAINERS: they initialise parameters (PyTorch's default nn.Linear / PReLU /
BatchNorm1d initialisers, as the reference gets them), carry them into and out of the fused kernel's
state block (engine.load_modules / store_modules) and are what `final.pt` pickles
(trainer.py:281-283, 310).  Training never calls their forward(); it exists so that a trained model
can be evaluated by the reference's report tooling (`sc/report/analysis.py:394-450`).
"""
import torch
from torch import nn
from torch.autograd import Function

HIDDEN = 64   # hard-coded default hidden_size in the reference (model.py:342, 528, 632)


class GradientReversalLayer(Function):
    """Identity forward; backward multiplies by -beta (no-op for beta None)."""

    @staticmethod
    def forward(ctx, x, beta):
        ctx.beta = beta
        return x

    @staticmethod
    def backward(ctx, grad_output):
        g = grad_output.clone()
        if ctx.beta is not None:
            g = -g * ctx.beta
        return g, None


def _hidden_block(n_in, n_out, dropout_rate, batch_norm=True):
    layers = [nn.Linear(n_in, n_out), nn.PReLU(num_parameters=n_out, init=0.01)]
    if batch_norm:
        layers.append(nn.BatchNorm1d(n_out, affine=False))
    layers.append(nn.Dropout(p=dropout_rate))
    return layers


class FCEncoder(nn.Module):
    def __init__(self, dropout_rate=0.2, nstyle=5, dim_in=256, n_layers=3, hidden_size=HIDDEN):
        super().__init__()
        if hidden_size != HIDDEN:
            raise ValueError("the fused path supports hidden_size == 64 only")
        seq = _hidden_block(dim_in, hidden_size, dropout_rate)
        for _ in range(n_layers - 2):
            seq += _hidden_block(hidden_size, hidden_size, dropout_rate)
        seq += [nn.Linear(hidden_size, nstyle), nn.BatchNorm1d(nstyle, affine=False)]
        self.main = nn.Sequential(*seq)
        self._ctor = dict(dropout_rate=dropout_rate, nstyle=nstyle, dim_in=dim_in, n_layers=n_layers)

    def forward(self, spec):
        return self.main(spec)


class FCDecoder(nn.Module):
    def __init__(self, dropout_rate=0.2, nstyle=5, debug=False, dim_out=256, last_layer_activation="ReLu",
                 n_layers=3, hidden_size=HIDDEN):
        super().__init__()
        if hidden_size != HIDDEN:
            raise ValueError("the fused path supports hidden_size == 64 only")
        if last_layer_activation == "ReLu":
            ll_act = nn.ReLU()
        elif last_layer_activation == "Softplus":
            ll_act = nn.Softplus(beta=2)
        else:
            raise ValueError(
                f"Unknow activation function \"{last_layer_activation}\", please use one available in Pytorch")
        seq = _hidden_block(nstyle, hidden_size, dropout_rate)
        for _ in range(n_layers - 2):
            seq += _hidden_block(hidden_size, hidden_size, dropout_rate)
        seq += [nn.Linear(hidden_size, dim_out), ll_act]
        self.main = nn.Sequential(*seq)
        self.nstyle = nstyle
        self.debug = debug
        self._ctor = dict(dropout_rate=dropout_rate, nstyle=nstyle, dim_out=dim_out, last_layer_activation=last_layer_activation,
                          n_layers=n_layers)

    def forward(self, z_gauss):
        return self.main(z_gauss)


class DiscriminatorFC(nn.Module):
    def __init__(self, hiden_size=HIDDEN, dropout_rate=0.2, nstyle=5, noise=0.1, layers=3):
        super().__init__()
        if hiden_size != HIDDEN:
            raise ValueError("the fused path supports hiden_size == 64 only")
        seq = _hidden_block(nstyle, hiden_size, dropout_rate, batch_norm=False)
        for _ in range(layers - 2):
            seq += _hidden_block(hiden_size, hiden_size, dropout_rate, batch_norm=False)
        seq += [nn.Linear(hiden_size, 1)]
        self.main = nn.Sequential(*seq)
        self.nstyle = nstyle
        self.noise = noise
        self._ctor = dict(dropout_rate=dropout_rate, nstyle=nstyle, noise=noise, layers=layers)

    def forward(self, x, beta):
        if self.training:
            x = x + self.noise * torch.randn_like(x, requires_grad=False)
        return self.main(GradientReversalLayer.apply(x, beta))
