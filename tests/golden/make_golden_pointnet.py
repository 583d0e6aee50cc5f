"""Golden vectors for the PointNet2Stage producer, from the REFERENCE's own module (build container only).

    python tests/golden/make_golden_pointnet.py    # needs /root/reference; writes tests/golden/pointnet_*.npz

Runs ``models.pointnet.PointNet2Stage(6, 128, 32).forward`` (reference src/models/pointnet.py:7-38, unmodified;
torch_scatter.scatter(reduce='max') stubbed as in make_golden.py) on seeded points with the reference's own random
initialisation, and stores weights, inputs and the output.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as MG  # noqa: E402


def run(name, seed, N, V, empty_voxels=()):
    MG.build_reference({})                                          # installs the stubs, puts the reference on sys.path
    import models.pointnet as ref_pointnet
    torch.manual_seed(seed)
    net = ref_pointnet.PointNet2Stage(input_channels=6, output_channels=128, gf_dim=32).eval()
    g = torch.Generator().manual_seed(seed + 1)
    inp = torch.cat((0.25 * (torch.rand(N, 3, generator=g) - 0.5), torch.rand(N, 3, generator=g)), 1)   # rel coord | rgb
    idx = torch.randint(0, V, (N,), generator=g)
    for e in empty_voxels:                                          # voxels without any point keep a zero feature
        idx[idx == e] = (e + 1) % V
    idx[0] = V - 1                                                  # the output has V rows (index.max() + 1)
    with torch.no_grad():
        out = net(inp_feat=inp, vox2point_idx=idx)
    save = {"inp_feat": inp.numpy(), "vox2point_idx": idx.numpy().astype(np.int32), "V": V, "ref_out": out.numpy()}
    # gradients of sum(c * out) through the reference module (its scatter-max stub routes the gradient to the arg-max rows
    # exactly like torch_scatter does: ties are not expected with continuous random inputs)
    c = torch.randn(out.shape, generator=g)
    inp_g = inp.clone().requires_grad_(True)
    net.train()
    out_g = net(inp_feat=inp_g, vox2point_idx=idx)
    (c * out_g).sum().backward()
    save["c_out"] = c.numpy(); save["grad.inp_feat"] = inp_g.grad.numpy()
    for k, p_ in net.named_parameters():
        save["grad." + k] = p_.grad.numpy()
    for k, v in net.state_dict().items():
        save["w." + k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **save)
    print(name, "N", N, "V", V, "out", tuple(out.shape), "nonzero rows", int((out.abs().sum(1) > 0).sum()))


if __name__ == "__main__":
    run("pointnet_2000x60", 31, 2000, 60, empty_voxels=(7, 33))
    run("pointnet_257x3", 32, 257, 3)
