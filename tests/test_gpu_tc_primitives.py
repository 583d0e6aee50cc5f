"""GPU: unit test of the tcgen05 building blocks of the decoder engine (include/lidf_query.h: lidf_tc_selftest)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tcgen05_split_bf16_mma_matches_fp32():
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 32, generator=g).cuda()
    W = (torch.randn(128, 32, generator=g) / 32 ** 0.5).cuda()
    want = (A.double() @ W.double().t()).float()
    got = lidf_query.tc_selftest(A, W, 0)
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 1e-4, err                       # 3-product bf16 split: ~1e-5; a wrong layout gives O(1)
    # exactly representable operands -> exact result (checks element order inside the packed words)
    Ai = torch.randint(-8, 9, (128, 32), generator=g).float().cuda()
    Wi = torch.randint(-8, 9, (128, 32), generator=g).float().cuda()
    assert torch.equal(lidf_query.tc_selftest(Ai, Wi, 0), Ai @ Wi.t())
