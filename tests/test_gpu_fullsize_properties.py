"""Size-independent properties at BASELINE.json's full shapes (640x480 views, 2048^2 x 4-layer texture, VGG up to
conv5_1), where the CPU oracle would take minutes: symmetry / trace identities of the Gram, linearity of the data
gradient, zero-mask and all-ones-mask limits, Adam no-op on zero gradient, replica determinism."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

H, W = 480, 640


@pytest.fixture(scope="module")
def engine():
    from stylemesh_b200 import engine as eng
    from stylemesh_b200 import synthetic as syn
    e = eng.VGGEngine(syn.make_vgg_state_dict(0, bias_scale=0.05))
    yield e
    e.close()


def _img(seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(3, H, W, generator=g) * 255 - 120).cuda()


def test_vgg_forward_shapes_and_gram_identities(engine):
    slot = engine.begin(H, W)
    engine.forward(slot, _img(0), 12)
    shapes = {0: (64, 480, 640), 2: (128, 240, 320), 4: (256, 120, 160), 8: (512, 60, 80), 12: (512, 30, 40)}
    for conv, shp in shapes.items():
        assert engine.feature_shape(slot, conv) == shp
        c, h, w = shp
        G = engine.gram(slot, conv, None, 1.0 / (h * w))
        assert torch.isfinite(G).all()
        assert torch.allclose(G, G.t(), rtol=1e-5, atol=1e-5 * float(G.abs().max()))           # symmetry
        f = engine.feature(slot, conv)
        trace = (f.double() ** 2).sum() / (h * w)                                               # tr G = |F|^2 / N
        assert abs(float(G.double().trace()) - float(trace)) <= 2e-5 * float(trace)
        assert float(f.min()) >= 0.0                                                            # post-ReLU


def test_masked_gram_limits(engine):
    slot = engine.begin(H, W)
    engine.forward(slot, _img(1), 4)
    c, h, w = engine.feature_shape(slot, 4)
    ones = torch.ones(h * w, device="cuda")
    zeros = torch.zeros(h * w, device="cuda")
    G_none = engine.gram(slot, 4, None, 1.0 / (h * w))
    G_ones = engine.gram(slot, 4, ones, 1.0 / (h * w))
    assert torch.equal(G_none, G_ones)                       # an all-ones mask is the identity, bit for bit
    assert float(engine.gram(slot, 4, zeros, 0.0).abs().max()) == 0.0
    half = ones.clone(); half[: (h * w) // 2] = 0
    G_a = engine.gram(slot, 4, half, 1.0)
    G_b = engine.gram(slot, 4, 1 - half, 1.0)
    G_all = engine.gram(slot, 4, None, 1.0)
    assert (G_a + G_b - G_all).norm() <= 1e-5 * G_all.norm()  # disjoint masks partition the Gram


def test_backward_is_linear_in_the_loss_weight_and_deterministic(engine):
    """d(pred) for content coefficient 2c equals 2 x d(pred) for c; two identical evaluations are bit-identical
    (fixed-order split reductions, stream-K partial sums in fixed order)."""
    img = _img(2)
    tgt_slot = engine.begin(H, W)
    engine.forward(tgt_slot, _img(3), 9)
    T = engine.feature_nhwc(tgt_slot, 9).clone()
    c, h, w = engine.feature_shape(tgt_slot, 9)
    mask = torch.ones(h * w, device="cuda")
    outs = []
    for coef in (1e-3, 2e-3, 1e-3):
        slot = engine.begin(H, W)
        engine.forward(slot, img, 9)
        acc = torch.zeros(1, device="cuda")
        engine.content_term(slot, 9, T, mask, coef, 2 * coef, acc)
        outs.append((engine.backward(slot, H, W).clone(), float(acc)))
    (g1, l1), (g2, l2), (g3, l3) = outs
    assert torch.equal(g1, g3)                      # the gradient path has no atomics: bit-identical replays
    assert abs(l1 - l3) <= 1e-6 * abs(l1)           # the scalar loss is a float-atomic sum over blocks
    assert abs(l2 - 2 * l1) <= 1e-5 * abs(l2)
    assert (g2 - 2 * g1).norm() <= 1e-4 * g2.norm()
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0


def test_adam_leaves_untouched_zero_texels_at_zero_and_4096_texture_fits():
    """reference: texels never touched by a view with zero init stay exactly 0 (m = v = 0 -> 0/(0+eps)); run on the
    4096^2 texture of BASELINE config C4 (201 MB per buffer)."""
    from stylemesh_b200 import engine as eng
    n = 3 * 4096 * 4096
    p = torch.zeros(n, device="cuda"); g = torch.zeros(n, device="cuda")
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    g[:1000] = 1.0
    eng.adam_step(p, g, m, v, 1.0, 0.9, 0.999, 1e-8, 1)
    assert float(p[1000:].abs().max()) == 0.0 and float(g.abs().max()) == 0.0
    assert torch.allclose(p[:1000], torch.full((1000,), -1.0, device="cuda"), atol=1e-6)
