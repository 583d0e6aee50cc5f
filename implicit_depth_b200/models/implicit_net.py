"""Mirror of the reference's ``models.implicit_net`` surface (reference src/models/implicit_net.py).

Same names, constructor signatures, attribute names and ``state_dict`` keys (``linear_{1..4}.{weight,bias}``,
``offset_enc.*``), so reference checkpoints load unchanged (src/trainers/train_lidf.py:90-112,349-371) and the
trainers' optimiser set-up keeps working (train_lidf.py:61-62).

The modules are parameter containers for the fused ``lidf_query`` kernel: ``LIDF.get_pred`` hands their tensors to the
C ABI and never calls ``forward``.  ``forward`` itself is kept for stand-alone / autograd use and is written with
ordinary torch ops (it is what DDP training differentiates through until a native backward exists).
Unlike the reference, importing this module does NOT flip ``torch.autograd.set_detect_anomaly(True)``
(implicit_net.py:2), a debugging switch that slows every backward pass.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Embedder:
    """NeRF positional encoding, reference implicit_net.py:9-39."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.create_embedding_fn()

    def create_embedding_fn(self):
        d = self.kwargs['input_dims']
        n_freqs = self.kwargs['num_freqs']
        max_freq = self.kwargs['max_freq_log2']
        if self.kwargs['log_sampling']:
            self.freq_bands = 2. ** torch.linspace(0., max_freq, steps=n_freqs)
        else:
            self.freq_bands = torch.linspace(2. ** 0., 2. ** max_freq, steps=n_freqs)
        self.periodic_fns = list(self.kwargs['periodic_fns'])
        self.include_input = bool(self.kwargs['include_input'])
        self.out_dim = (d if self.include_input else 0) + d * len(self.periodic_fns) * n_freqs

    def embed(self, inputs):
        parts = [inputs] if self.include_input else []
        for freq in self.freq_bands.tolist():
            for p_fn in self.periodic_fns:
                parts.append(p_fn(inputs * freq))
        return torch.cat(parts, -1)


def get_embedder(multires, i=0):
    """Reference implicit_net.py:42-57: returns (callable, out_dim); ``i == -1`` disables the encoding."""
    if i == -1:
        return nn.Identity(), 3
    embedder_obj = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                            log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    embed = lambda x, eo=embedder_obj: eo.embed(x)
    embed.multires = multires
    return embed, embedder_obj.out_dim


def _init_mlp(mod):
    """Reference implicit_net.py:72-79 / :117-126."""
    for i in (1, 2, 3):
        lin = getattr(mod, f'linear_{i}')
        nn.init.normal_(lin.weight, mean=0.0, std=0.02)
        nn.init.constant_(lin.bias, 0)
    nn.init.normal_(mod.linear_4.weight, mean=1e-5, std=0.02)
    nn.init.constant_(mod.linear_4.bias, 0)


def _final_act(x, use_sigmoid):
    if use_sigmoid:
        return torch.sigmoid(x)
    return torch.max(torch.min(x, x * 0.01 + 0.99), x * 0.01)


class IMNet(nn.Module):
    """Reference implicit_net.py:60-98."""

    def __init__(self, inp_dim, out_dim, gf_dim=64, use_sigmoid=False):
        super(IMNet, self).__init__()
        self.inp_dim = inp_dim
        self.gf_dim = gf_dim
        self.use_sigmoid = use_sigmoid
        self.linear_1 = nn.Linear(self.inp_dim, self.gf_dim * 4, bias=True)
        self.linear_2 = nn.Linear(self.gf_dim * 4, self.gf_dim * 2, bias=True)
        self.linear_3 = nn.Linear(self.gf_dim * 2, self.gf_dim * 1, bias=True)
        self.linear_4 = nn.Linear(self.gf_dim * 1, out_dim, bias=True)
        if self.use_sigmoid:
            self.sigmoid = nn.Sigmoid()
        _init_mlp(self)

    def forward(self, inp_feat):
        h = F.leaky_relu(self.linear_1(inp_feat), negative_slope=0.02)
        h = F.leaky_relu(self.linear_2(h), negative_slope=0.02)
        h = F.leaky_relu(self.linear_3(h), negative_slope=0.02)
        return _final_act(self.linear_4(h), self.use_sigmoid)


class IEF(nn.Module):
    """Reference implicit_net.py:100-152 (iterative error feedback decoder)."""

    def __init__(self, device, inp_dim, out_dim, gf_dim=64, n_iter=3, use_sigmoid=False):
        super(IEF, self).__init__()
        self.device = device
        self.init_offset = torch.Tensor([0.001]).float().to(self.device)   # plain attribute, not a buffer (:104)
        self.inp_dim = inp_dim
        self.gf_dim = gf_dim
        self.n_iter = n_iter
        self.use_sigmoid = use_sigmoid
        self.offset_enc = nn.Linear(1, 16, bias=True)
        self.linear_1 = nn.Linear(self.inp_dim + 16, self.gf_dim * 4, bias=True)
        self.linear_2 = nn.Linear(self.gf_dim * 4, self.gf_dim * 2, bias=True)
        self.linear_3 = nn.Linear(self.gf_dim * 2, self.gf_dim * 1, bias=True)
        self.linear_4 = nn.Linear(self.gf_dim * 1, out_dim, bias=True)
        if self.use_sigmoid:
            self.sigmoid = nn.Sigmoid()
        nn.init.normal_(self.offset_enc.weight, mean=0.0, std=0.02)
        nn.init.constant_(self.offset_enc.bias, 0)
        _init_mlp(self)

    def forward(self, inp_feat):
        pred_offset = self.init_offset.to(inp_feat.device).expand(inp_feat.shape[0], -1)
        for _ in range(self.n_iter):
            xc = torch.cat([inp_feat, self.offset_enc(pred_offset)], 1)
            h = F.leaky_relu(self.linear_1(xc), negative_slope=0.02)
            h = F.leaky_relu(self.linear_2(h), negative_slope=0.02)
            h = F.leaky_relu(self.linear_3(h), negative_slope=0.02)
            pred_offset = pred_offset + self.linear_4(h)
        return _final_act(pred_offset, self.use_sigmoid)
