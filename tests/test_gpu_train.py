"""GPU parity of the training-side front-end kernels (weight gradient on the tensor cores) against torch fp32 autograd
on the same bf16-rounded inputs."""
import numpy as np
import pytest
import torch

from doubleattentionspeakerverification_b200 import ops

pytestmark = pytest.mark.gpu


def ref_wgrad(x, g):
    """torch fp32 reference: d/dW of conv2d(x, W, padding=1) contracted with g.  x [B,T,F,Cin], g [B,T,F,Cout] (NHWC)."""
    xr = x.float().permute(0, 3, 1, 2).contiguous()
    gr = g.float().permute(0, 3, 1, 2).contiguous()
    w = torch.zeros((g.shape[3], x.shape[3], 3, 3), device=x.device, requires_grad=True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = torch.nn.functional.conv2d(xr, w, padding=1)
        (y * gr).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return w.grad


@pytest.mark.parametrize('B,T,F,Cin,Cout', [(1, 8, 10, 64, 128), (2, 13, 20, 128, 128), (3, 50, 10, 64, 256),
                                            (2, 9, 40, 128, 256), (1, 6, 80, 128, 128), (5, 7, 6, 192, 128)])
def test_wgrad_vs_autograd(B, T, F, Cin, Cout):
    gen = torch.Generator(device='cuda').manual_seed(B * 100 + T)
    x = torch.randn(B, T, F, Cin, device='cuda', generator=gen).to(torch.bfloat16)
    g = (torch.randn(B, T, F, Cout, device='cuda', generator=gen) * 0.5).to(torch.bfloat16)
    dw = ops.conv3x3_wgrad(x, g)
    want = ref_wgrad(x, g)
    err = float((dw - want).abs().max() / want.abs().max())
    assert err < 2e-4, err
    dw2 = ops.conv3x3_wgrad(x, g, dw.clone())                    # accumulate
    assert float((dw2 - 2 * want).abs().max() / want.abs().max()) < 4e-4
    assert torch.equal(ops.conv3x3_wgrad(x, g), dw)              # deterministic
