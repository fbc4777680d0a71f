"""GPU parity of single VGG-side kernels through the unit C-ABI entry points, for BOTH implementations:
SIMT (fp32 CUDA cores) and TC (tcgen05 tensor cores, bf16x3 split) — against torch fp32 CPU ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

IMPLS = [pytest.param(0, id="simt"), pytest.param(1, id="tc"), pytest.param(5, id="ph")]
GRAM_IMPLS = [pytest.param(0, id="simt"), pytest.param(1, id="tc")]


def _eng():
    from stylemesh_b200 import engine
    return engine


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


CONV_SHAPES = [(64, 64, 24, 32), (64, 128, 17, 23), (128, 256, 12, 17), (256, 256, 9, 12), (512, 512, 6, 8),
               (256, 512, 3, 4), (64, 256, 120, 160),   # 150 M-tiles x 256-wide N tile (v1: BN=256)
               (512, 512, 30, 40), (256, 256, 60, 80), (128, 128, 33, 47)]   # stream-K: tiles split over many CTAs


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("cin,cout,h,w", CONV_SHAPES)
def test_conv3x3_bias_relu_forward(impl, cin, cout, h, w):
    eng = _eng()
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(cin, h, w, generator=g) * 50
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(x.unsqueeze(0), wt, b, padding=1))[0]
    out = eng.unit_conv3x3(impl, x.cuda(), wt, b, relu=True).cpu()
    assert _rel(out, ref) < 5e-5, _rel(out, ref)
    assert torch.allclose(out, ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("cin,cout,h,w", CONV_SHAPES)
def test_conv3x3_data_gradient(impl, cin, cout, h, w):
    eng = _eng()
    g = torch.Generator().manual_seed(7 * cin + cout + w)
    x = torch.randn(1, cin, h, w, generator=g).requires_grad_(True)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    dy = torch.randn(1, cout, h, w, generator=g)
    F.conv2d(x, wt, None, padding=1).backward(dy)
    out = eng.unit_conv3x3(impl, dy[0].cuda(), wt, None, relu=False, transpose_flip=True).cpu()
    assert _rel(out, x.grad[0]) < 5e-5, _rel(out, x.grad[0])


@pytest.mark.parametrize("masked", [False, True], ids=["nomask", "mask"])
@pytest.mark.parametrize("cin,cout,h,w", [(64, 64, 32, 40), (128, 128, 24, 32), (64, 128, 17, 23), (256, 256, 16, 16),
                                          (512, 512, 30, 40), (64, 64, 120, 160), (128, 64, 33, 47)])
def test_ph_conv_with_fused_gram_backward_term(cin, cout, h, w, masked):
    """conv3x3(x) + m * (g f): the Gram backward (cs:74-80 through torch.bmm's backward) folded into the pair + halo
    data-gradient conv as extra K-chunks; the mask zeroes feature pixels in shared memory."""
    eng = _eng()
    gen = torch.Generator().manual_seed(3 * cin + cout + h + int(masked))
    x = torch.randn(cin, h, w, generator=gen) * 5
    wt = torch.randn(cout, cin, 3, 3, generator=gen) * (2.0 / (9 * cin)) ** 0.5
    f = F.relu(torch.randn(cout, h, w, generator=gen)) * 20
    gm = torch.randn(cout, cout, generator=gen) / cout
    gm = 0.5 * (gm + gm.t())
    mask = (torch.rand(h * w, generator=gen) > 0.35).float() if masked else None
    term = torch.einsum("nk,khw->nhw", gm.double(), f.double())
    if masked:
        term = term * mask.reshape(1, h, w).double()
    ref = F.conv2d(x.unsqueeze(0).double(), wt.double(), None, padding=1)[0] + term
    out = eng.unit_conv3x3_fused(x.cuda(), wt, f.cuda(), gm, None if mask is None else mask.cuda()).cpu()
    assert _rel(out.double(), ref) < 5e-5, _rel(out.double(), ref)
    if masked:      # masked pixels carry the plain convolution only
        conv_only = F.conv2d(x.unsqueeze(0).double(), wt.double(), None, padding=1)[0]
        sel = (mask == 0).reshape(h, w)
        assert _rel(out.double()[:, sel], conv_only[:, sel]) < 5e-5


@pytest.mark.parametrize("h,w", [(24, 32), (17, 23), (2, 2), (5, 3)])
def test_maxpool_forward_backward(h, w):
    eng = _eng()
    g = torch.Generator().manual_seed(h * w)
    y = F.relu(torch.randn(1, 64, h, w, generator=g)).requires_grad_(True)     # post-ReLU (ties at 0 happen)
    p = F.max_pool2d(y, 2, 2)
    gp = torch.randn(p.shape, generator=g)
    p.backward(gp)
    out = eng.unit_maxpool(y.detach()[0].cuda()).cpu()
    assert torch.allclose(out, p.detach()[0], rtol=1e-5, atol=1e-6)
    # our backward also applies the ReLU mask of y (y > 0)
    want = y.grad[0] * (y.detach()[0] > 0)
    dx = eng.unit_maxpool_bwd(gp[0].cuda(), y.detach()[0].cuda()).cpu()
    assert torch.allclose(dx, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("impl", GRAM_IMPLS)
@pytest.mark.parametrize("c,h,w,masked", [(64, 24, 32, False), (64, 48, 64, True), (128, 24, 32, True),
                                          (256, 12, 16, False), (512, 6, 8, True), (512, 3, 4, False),
                                          (256, 31, 37, True), (64, 120, 160, True), (512, 60, 80, True),
                                          (128, 97, 131, True)])
def test_masked_gram(impl, c, h, w, masked):
    eng = _eng()
    g = torch.Generator().manual_seed(c + h + w)
    f = F.relu(torch.randn(c, h, w, generator=g)) * 30
    mask = (torch.rand(h * w, generator=g) > 0.3).float() if masked else None
    fm = f.reshape(c, -1) * (mask if masked else 1.0)
    n = float(mask.sum()) if masked else float(h * w)
    ref = fm.double() @ fm.double().t() / n
    out = eng.unit_gram(impl, f.cuda(), None if mask is None else mask.cuda(), 1.0 / n).cpu()
    assert _rel(out.double(), ref) < 2e-5, _rel(out.double(), ref)
    assert torch.allclose(out, out.t(), rtol=1e-5, atol=1e-5 * float(out.abs().max()))     # symmetry property


@pytest.mark.parametrize("h,w", [(37, 53), (64, 128), (48, 200), (17, 16)])
def test_first_layer_tcgen05_ragged_sizes(h, w, tmp_path):
    """conv1_1 (3 -> 64, software im2col + TMA-store epilogue): image sizes that do not divide the 128-pixel patch, so
    the tensor store has to clip rows / columns (cs:49 relu(conv1_1(x)))."""
    from stylemesh_b200 import synthetic as syn
    from stylemesh_b200.model.losses.content_and_style_losses import VGG
    sd = syn.make_vgg_state_dict(3, bias_scale=0.5)
    path = str(tmp_path / "vgg.pth")
    torch.save(sd, path)
    vgg = VGG(model_path=path).cuda()
    g = torch.Generator().manual_seed(h * w)
    x = torch.rand(1, 3, h, w, generator=g) * 255 - 110
    got = vgg(x.cuda(), ["r11"])["r11"].cpu()
    want = F.relu(F.conv2d(x, sd["conv1_1.weight"], sd["conv1_1.bias"], padding=1))
    assert got.shape == want.shape
    assert _rel(got, want) < 2e-5, _rel(got, want)


@pytest.mark.parametrize("h,w", [(64, 96), (37, 53), (48, 200), (33, 64)])
def test_inference_only_forward_pools_in_the_conv_epilogue(h, w, tmp_path):
    """smb_level_forward_features (the content-target pass, cs:294): layers that only feed a pool leave igemm_ph through
    the 2x2-max epilogue.  Max-pooling selects one of four values, so the kept features are value-identical to the
    training forward's (odd sizes: MaxPool2d floor mode drops the trailing row / column)."""
    from stylemesh_b200 import engine as E, synthetic as syn
    sd = syn.make_vgg_state_dict(5, bias_scale=0.3)
    eng = E.VGGEngine(sd)
    g = torch.Generator().manual_seed(h + w)
    img = (torch.rand(3, h, w, generator=g) * 255 - 110).cuda()
    slot = eng.begin(h, w)
    eng.forward(slot, img, 9)
    want = {c: eng.feature(slot, c).clone() for c in (2, 4, 8, 9)}
    eng.forward(slot, img, 9, keep=[9])
    assert torch.equal(eng.feature(slot, 9), want[9])
    from stylemesh_b200._abi import StyleMeshB200Error
    with pytest.raises(StyleMeshB200Error, match="not kept"):
        eng.feature(slot, 3)                      # conv2_2 only fed pool2: never materialised
    eng.forward(slot, img, 9, keep=[2, 4, 8, 9])  # conv2_1 / conv3_1 / conv4_1 follow the pools: always materialised
    for c in (2, 4, 8, 9):
        assert torch.equal(eng.feature(slot, c), want[c]), c
    eng.forward(slot, img, 9, keep=[1, 3, 7, 9])  # keeping the pool feeders falls back to the separate pool kernel
    assert torch.equal(eng.feature(slot, 9), want[9])
    eng.forward(slot, img, 9)                     # a training forward afterwards sees nothing stale
    for c in (2, 4, 8, 9):
        assert torch.equal(eng.feature(slot, c), want[c]), c
