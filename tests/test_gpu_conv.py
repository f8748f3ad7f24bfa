"""GPU parity of the general convolution operator (pmvae_conv2d_forward / backward) against oracle/conv.py for
every layer of the MNIST config's ConvEncoder / ConvDecoder, forward and all three gradients."""
import numpy as np
import pytest
import torch

from oracle import conv as OC
from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _layer_cases():
    cases = []
    H, cin = 28, 2                                     # partial encoder input [x*b, b]
    for i, (f, k, s) in enumerate(OC.MNIST_ENCODER):
        pad = "VALID" if i == len(OC.MNIST_ENCODER) - 1 else "SAME"
        cases.append(("enc%d" % i, False, H, cin, f, k, s, pad))
        H = (H - k) // s + 1 if pad == "VALID" else -(-H // s)
        cin = f
    H, cin = 1, 32
    for i, (f, k, s) in enumerate(OC.MNIST_DECODER):
        pad = "VALID" if i == 0 else "SAME"
        cases.append(("dec%d" % i, True, H, cin, f, k, s, pad))
        H = H * s + (max(k - s, 0) if pad == "VALID" else 0)
        cin = f
    return cases


@pytest.mark.parametrize("direct", [False, True, "bf16"], ids=["im2col", "direct", "bf16"])
@pytest.mark.parametrize("name,transpose,H,cin,cout,k,s,pad", _layer_cases())
def test_conv_layer_forward_and_gradients(name, transpose, H, cin, cout, k, s, pad, direct):
    from posterior_matching_b200 import conv as PC
    torch.manual_seed(hash(name) % 1000)
    B = 3
    x = torch.randn(B, H, H, cin, dtype=torch.float64, requires_grad=True)
    wshape = (k, k, cout, cin) if transpose else (k, k, cin, cout)
    w = (torch.randn(wshape, dtype=torch.float64) / (k * cin ** 0.5)).requires_grad_(True)
    b = (0.1 * torch.randn(cout, dtype=torch.float64)).requires_grad_(True)
    want = (OC.conv2d_transpose if transpose else OC.conv2d)(x, w, b, s, pad)
    g = torch.randn_like(want)
    (want * g).sum().backward()

    bf16 = direct == "bf16"
    direct = direct is True
    d = PC.conv_desc(H, H, cin, cout, k, s, pad, transpose=transpose, precision="bf16" if bf16 else "fp32")
    assert (d.OH, d.OW) == tuple(want.shape[1:3])
    xc, wc, bc = (t.detach().float().cuda().contiguous() for t in (x, w, b))
    y = PC.conv2d_forward(d, xc, wc, bc, direct=direct)
    dw, db = torch.zeros_like(wc), torch.zeros_like(bc)
    # bf16: the VJP is checked with the oracle's activations, so that sign flips of near-zero pre-activations
    # (a property of the forward rounding, covered by the model-level tolerance) do not mask GEMM errors
    y_bwd = want.detach().float().cuda().contiguous() if bf16 else y
    dx = PC.conv2d_backward(d, xc, wc, y_bwd, g.float().cuda().contiguous(), dw, db, direct=direct)
    torch.cuda.synchronize()
    if bf16:
        # bf16 GEMM operands, fp32 accumulation: operand rounding is 2^-9 relative per element
        assert rel_l2(y.cpu().numpy(), want.detach().numpy()) < 6e-3
        assert rel_l2(dx.cpu().numpy(), x.grad.numpy()) < 1e-2
        assert rel_l2(dw.cpu().numpy(), w.grad.numpy()) < 1e-2
        assert rel_l2(db.cpu().numpy(), b.grad.numpy()) < 5e-5
        return
    assert rel_err(y.cpu().numpy(), want.detach().numpy()) < 2e-5
    assert rel_l2(dx.cpu().numpy(), x.grad.numpy()) < 5e-5
    assert rel_l2(dw.cpu().numpy(), w.grad.numpy()) < 5e-5
    assert rel_l2(db.cpu().numpy(), b.grad.numpy()) < 5e-5
