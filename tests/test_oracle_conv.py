"""Pins oracle/conv.py: output geometry of the MNIST networks (configs/pm_vae_mnist.py) and agreement of the
explicit zero-insertion transposed convolution with torch's own conv_transpose2d where the two conventions
coincide (VALID, kernel flipped)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import conv as OC


def _params(layers, cin, transpose, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for f, k, _ in layers:
        shape = (k, k, f, cin) if transpose else (k, k, cin, f)
        out.append((torch.randn(shape, generator=g, dtype=torch.float64) / (k * cin ** 0.5),
                    0.1 * torch.randn(f, generator=g, dtype=torch.float64)))
        cin = f
    return out


def test_mnist_network_geometry():
    x = torch.randn(2, 28, 28, 1, dtype=torch.float64)
    h = OC.conv_encoder(_params(OC.MNIST_ENCODER, 1, False), x)
    assert tuple(h.shape) == (2, 1, 1, 128)              # 28 -> 28 -> 14 -> 14 -> 7 -> 1 (7x7 VALID)
    z = torch.randn(2, 32, dtype=torch.float64)
    y = OC.conv_decoder(_params(OC.MNIST_DECODER, 32, True), z)
    assert tuple(y.shape) == (2, 28, 28, 1)              # 1 -> 7 -> 14 -> 14 -> 28 -> 28 -> 28


def test_valid_transpose_equals_torch_conv_transpose_with_flipped_kernel():
    torch.manual_seed(1)
    x = torch.randn(3, 4, 5, 6, dtype=torch.float64)
    w = torch.randn(3, 3, 7, 6, dtype=torch.float64)      # [kh,kw,O,I]
    b = torch.randn(7, dtype=torch.float64)
    for s in (1, 2):
        got = OC.conv2d_transpose(x, w, b, s, "VALID", slope=1.0)
        # torch's conv_transpose2d scatters x[i] * W: equal to a correlation of the dilated input with the flipped kernel
        wt = torch.flip(w, dims=(0, 1)).permute(3, 2, 0, 1)            # [I,O,kh,kw]
        want = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, b, stride=s).permute(0, 2, 3, 1)
        if s == 2:
            want = F.pad(want, [0, 0, 0, 1, 0, 1])        # lax VALID keeps max(k - s, 0) extra: (in-1)s + 1 + ... vs torch (in-1)s + k
            want = want[:, :got.shape[1], :got.shape[2]]
        assert got.shape[1] >= (x.shape[1] - 1) * s + 3
        n1, n2 = min(got.shape[1], want.shape[1]), min(got.shape[2], want.shape[2])
        assert torch.allclose(got[:, :n1, :n2], want[:, :n1, :n2], atol=1e-10)


def test_same_conv_matches_torch_same_padding_for_stride_one():
    torch.manual_seed(2)
    x = torch.randn(2, 9, 9, 3, dtype=torch.float64)
    w = torch.randn(5, 5, 3, 4, dtype=torch.float64)
    b = torch.randn(4, dtype=torch.float64)
    got = OC.conv2d(x, w, b, 1, "SAME", slope=1.0)
    want = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding="same").permute(0, 2, 3, 1)
    assert torch.allclose(got, want, atol=1e-12)
    assert tuple(OC.conv2d(x, w, b, 2, "SAME").shape) == (2, 5, 5, 4)


def _desc_dict(H, cin, cout, k, s, pad, transpose):
    from posterior_matching_b200 import conv as PC
    d = PC.conv_desc(H, H, cin, cout, k, s, pad, transpose=transpose)
    return {f: int(getattr(d, f)) for f in ("H", "W", "Cin", "OH", "OW", "Cout", "KH", "KW", "stride", "dil", "pad_top", "pad_left")}


_GENERAL_CASES = [(28, 2, 8, 5, 1, "SAME", False), (28, 8, 8, 5, 2, "SAME", False), (7, 8, 16, 7, 1, "VALID", False),
                  (1, 8, 8, 7, 1, "VALID", True), (7, 8, 8, 5, 2, "SAME", True), (14, 8, 1, 5, 1, "SAME", True)]


@pytest.mark.parametrize("H,cin,cout,k,s,pad,transpose", _GENERAL_CASES)
def test_general_operator_descriptor_reproduces_both_layer_types(H, cin, cout, k, s, pad, transpose):
    """The one-operator form both hk.Conv2D and hk.Conv2DTranspose are lowered to (descriptors from conv.py)."""
    torch.manual_seed(1)
    x = torch.randn(2, H, H, cin, dtype=torch.float64)
    w = torch.randn((k, k, cout, cin) if transpose else (k, k, cin, cout), dtype=torch.float64)
    d = _desc_dict(H, cin, cout, k, s, pad, transpose)
    taps = w.permute(0, 1, 3, 2) if transpose else w          # HWOI -> tap-major [kh, kw, ci, co]
    got = OC.general_conv(x, taps, d)
    want = (OC.conv2d_transpose if transpose else OC.conv2d)(x, w, None, s, pad, slope=1.0)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-10


@pytest.mark.parametrize("H,cin,cout,k,s,pad,transpose", _GENERAL_CASES)
def test_data_gradient_is_the_adjoint_convolution(H, cin, cout, k, s, pad, transpose):
    """dX = general_conv(dY, reversed taps with channels swapped, adjoint descriptor): the identity the bf16 path of
    csrc/conv.cu uses for its data gradient."""
    torch.manual_seed(2)
    x = torch.randn(2, H, H, cin, dtype=torch.float64, requires_grad=True)
    w = torch.randn((k, k, cout, cin) if transpose else (k, k, cin, cout), dtype=torch.float64)
    d = _desc_dict(H, cin, cout, k, s, pad, transpose)
    taps = w.permute(0, 1, 3, 2) if transpose else w
    y = OC.general_conv(x, taps, d)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    a = OC.adjoint_desc(d)
    assert a["pad_top"] >= 0 and a["pad_left"] >= 0
    taps_adj = torch.flip(taps, dims=(0, 1)).permute(0, 1, 3, 2)     # [kh, kw, co -> ci', ci -> co']
    dx = OC.general_conv(g, taps_adj, a)
    assert float((dx - x.grad).abs().max()) < 1e-10
