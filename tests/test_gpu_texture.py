"""GPU parity of the texture-side kernels (through the C-ABI) against the CPU oracle / torch fp32 CPU ops."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stylemesh_oracle as orc

pytestmark = pytest.mark.gpu


def _eng():
    from stylemesh_b200 import engine
    return engine


def _grid(seed, h, w):
    g = torch.Generator().manual_seed(seed)
    grid = torch.rand(h, w, 2, generator=g) * 2.2 - 1.1             # includes out-of-range coordinates
    grid[0, 0] = torch.tensor([-1.0, -1.0])                         # the "invalid pixel" value
    grid[0, 1] = torch.tensor([1.0, 1.0])
    grid[0, 2] = torch.tensor([1.0 - 1e-7, -1.0 + 1e-7])
    grid[0, 3] = torch.tensor([0.0, 0.0])
    return grid


@pytest.mark.parametrize("tex_wh", [(37, 23), (512, 512), (2048, 2048)])
def test_uv_index_math_bit_exact(tex_wh):
    """north_star: bit-exact UV index math — integer texels AND the fp32 weights equal the ATen formula."""
    eng = _eng()
    W, H = tex_wh
    grid = _grid(1, 61, 67)
    # exact texel centres as well
    xs = torch.arange(0, 8, dtype=torch.float32) / (W - 1) * 2 - 1
    grid[1, :8, 0] = xs
    x0, y0, w = orc.uv_texel_indices_np(grid.numpy(), W, H)
    xy_d, w_d = eng.uv_texel_index(grid.cuda(), W, H)
    xy_d = xy_d.cpu().numpy().reshape(61, 67, 2)
    assert np.array_equal(xy_d[..., 0], x0) and np.array_equal(xy_d[..., 1], y0)
    assert np.array_equal(w_d.cpu().numpy().reshape(61, 67, 4).view(np.uint32), w.view(np.uint32))


@pytest.mark.parametrize("num_layers", [1, 4])
def test_sample_forward_matches_grid_sample(num_layers):
    eng = _eng()
    g = torch.Generator().manual_seed(3)
    layers = [torch.rand(3, 96 // 2 ** i, 128 // 2 ** i, generator=g) * 300 - 150 for i in range(num_layers)]
    grid = _grid(2, 45, 53)
    ref = orc.texture_sample([l.clone() for l in layers], grid.unsqueeze(0))[0]      # clamps, then samples
    out = eng.uv_sample_fwd([l.cuda() for l in layers], grid.cuda()).cpu()
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-4)


def test_scatter_backward_matches_autograd_with_hooks():
    eng = _eng()
    g = torch.Generator().manual_seed(5)
    layers = [torch.rand(3, 64 // 2 ** i, 80 // 2 ** i, generator=g).requires_grad_(True) for i in range(3)]
    grid = _grid(4, 33, 47)
    gout = torch.randn(3, 33, 47, generator=g)
    hook0 = torch.rand(33, 47, generator=g)
    hook1 = torch.rand(33, 47, generator=g)
    out = orc.texture_sample(layers, grid.unsqueeze(0))
    out.backward((gout * hook0 * hook1).unsqueeze(0))
    grads = [torch.zeros_like(l).cuda() for l in layers]
    eng.uv_scatter_bwd(grads, grid.cuda(), gout.cuda(), hook0.cuda(), hook1.cuda())
    for gd, l in zip(grads, layers):
        assert (gd.cpu() - l.grad).norm() <= 1e-5 * l.grad.norm() + 1e-7


def test_full_size_partition_of_unity():
    """size-independent properties at the benchmark shape (2048^2 x 4 layers, 640x480 view)."""
    eng = _eng()
    from stylemesh_b200 import synthetic as syn
    view = syn.make_view(1000, (480, 640), [(480, 640)])
    grid = view.uvs[0][0].cuda()
    layers = [torch.full((3, 2048 // 2 ** i, 2048 // 2 ** i), float(i + 1), device="cuda") for i in range(4)]
    out = eng.uv_sample_fwd(layers, grid)
    assert torch.allclose(out, torch.full_like(out, 10.0), rtol=0, atol=1e-4)      # weights sum to 1 per layer
    gout = torch.rand(3, 480, 640, device="cuda")
    grads = [torch.zeros_like(l) for l in layers]
    eng.uv_scatter_bwd(grads, grid, gout, None, None)
    want = gout.double().sum(dim=(1, 2))
    for gl in grads:
        got = gl.double().sum(dim=(1, 2))
        assert torch.allclose(got, want, rtol=1e-4)


def test_fused_adam_clamp_reg_matches_torch():
    eng = _eng()
    g = torch.Generator().manual_seed(9)
    n = 3 * 33 * 35                                      # not a multiple of 4: exercises the scalar tail
    p0 = torch.rand(n, generator=g) * 400 - 200          # some values outside the clamp range
    grads = [torch.randn(n, generator=g) * (0.1 if k else 1.0) for k in range(4)]
    grads[1][::7] = 0.0
    lam_w = 5e3 * 4.0
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1.0)
    p, gd = p0.clone().cuda(), torch.zeros(n, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for k, gk in enumerate(grads):
        with torch.no_grad():
            ref.copy_(ref.clamp(orc.CLAMP_LO, orc.CLAMP_HI))
        opt.zero_grad()
        (lam_w * torch.mean(ref ** 2)).backward()
        ref.grad += gk
        opt.step()
        gd.copy_(gk.cuda() * 2.0)                        # pretend two ranks summed: grad_scale = 1/2
        eng.adam_step(p, gd, m, v, 1.0, 0.9, 0.999, 1e-8, k + 1, reg_coef=2.0 * lam_w / n, grad_scale=0.5)
        assert float(gd.abs().max()) == 0.0              # the kernel leaves the gradient buffer zeroed
        # lr = 1: the first update is +-1 * sign(g); compare with an absolute tolerance on O(100) values
        assert torch.allclose(p.cpu(), ref.detach(), rtol=0, atol=2e-4), k


def test_texreg_value():
    eng = _eng()
    x = torch.rand(3, 64, 64) * 400 - 200
    want = 7.0 * torch.mean(x.clamp(orc.CLAMP_LO, orc.CLAMP_HI) ** 2)
    acc = torch.zeros(1, device="cuda")
    eng.texreg_value(x.cuda(), 7.0 / x.numel(), acc)
    assert abs(float(acc) - float(want)) <= 1e-5 * float(want)


def test_segmented_adam_and_texreg_equal_the_per_layer_kernels():
    """One launch over the flat 4-layer buffer (64-float aligned segments with zero padding) == one launch per layer,
    bit for bit: same arithmetic per element, only the regulariser coefficient is looked up per segment."""
    eng = _eng()
    g = torch.Generator().manual_seed(21)
    sizes = [3 * 40 * 40, 3 * 20 * 20, 3 * 10 * 10, 3 * 5 * 5]
    begin, off = [], 0
    for n in sizes:
        begin.append(off)
        off += (n + 63) // 64 * 64
    coefs = [2.0 * 5e3 * w / n for w, n in zip([8.0, 4.0, 2.0, 0.0], sizes)]

    def fresh():
        p = torch.zeros(off)
        gr = torch.zeros(off)
        for a, n in zip(begin, sizes):
            p[a:a + n] = torch.rand(n, generator=torch.Generator().manual_seed(a)) * 400 - 200
            gr[a:a + n] = torch.randn(n, generator=torch.Generator().manual_seed(a + 1))
        return p.cuda(), gr.cuda(), torch.zeros(off, device="cuda"), torch.zeros(off, device="cuda")

    p1, g1, m1, v1 = fresh()
    p2, g2, m2, v2 = fresh()
    acc1, acc2 = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    for step in (1, 2, 3):
        eng.texreg_value_segments(p1, begin, [c / 2.0 for c in coefs], acc1)
        eng.adam_step_segments(p1, g1, m1, v1, begin, coefs, 1.0, 0.9, 0.999, 1e-8, step, grad_scale=0.5)
        for a, n, c in zip(begin, sizes, coefs):
            eng.texreg_value(p2[a:a + n], c / 2.0, acc2)
            eng.adam_step(p2[a:a + n], g2[a:a + n], m2[a:a + n], v2[a:a + n], 1.0, 0.9, 0.999, 1e-8, step,
                          reg_coef=c, grad_scale=0.5)
        assert torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)
        assert float(g1.abs().max()) == 0.0
        g1.copy_(torch.randn(off, generator=g).cuda() * (g2 * 0 + 1))
        g2.copy_(g1)
        for a, n in zip(begin, sizes):                     # padding carries no gradient
            g1[a + n:a + (n + 63) // 64 * 64] = 0
            g2[a + n:a + (n + 63) // 64 * 64] = 0
    for a, n in zip(begin, sizes):                         # padding stayed exactly zero
        assert float(p1[a + n:a + (n + 63) // 64 * 64].abs().max() if (n % 64) else 0.0) == 0.0
    assert abs(float(acc1) - float(acc2)) <= 1e-5 * abs(float(acc2))


# ---------------------------------------------------------------------------------------------------------------
# texture export and headless preview
# ---------------------------------------------------------------------------------------------------------------
def _reference_texture_bytes(layers):
    """texture.py:110-121 get_image + rgb_transform.py:14-21 post() + texture.py:9-19 ToPILImage, with the reference's
    torch ops on the CPU."""
    import torch.nn.functional as F
    from oracle import stylemesh_oracle as orc
    C, H, W = layers[0].shape
    w_range = torch.arange(0, W, dtype=torch.float) / (W - 1.0) * 2.0 - 1.0
    h_range = torch.arange(0, H, dtype=torch.float) / (H - 1.0) * 2.0 - 1.0
    v, u = torch.meshgrid(h_range, w_range, indexing="ij")
    img = orc.texture_sample([l.clone() for l in layers], torch.stack([u, v], 2).unsqueeze(0))[0, 0:3]
    x = img.clone().mul_(1.0 / 255)
    mean = torch.tensor([-0.40760392, -0.45795686, -0.48501961]).view(3, 1, 1)
    x = (x - mean) / 1.0
    x = x[torch.LongTensor([2, 1, 0])].clamp(0, 1)
    return x.mul(255).byte().permute(1, 2, 0).contiguous()


@pytest.mark.parametrize("size,layers", [(256, 4), (100, 3), (64, 1)])
def test_texture_export_bytes_match_the_reference_chain(size, layers, tmp_path):
    from stylemesh_b200 import export, synthetic as syn
    from stylemesh_b200.model.losses.rgb_transform import post
    from stylemesh_b200.model.texture.texture import HierarchicalNeuralTexture, NeuralTexture
    g = torch.Generator().manual_seed(size)
    ls = [(torch.rand(3, size // 2 ** i, size // 2 ** i, generator=g) * 300 - 140) / (i + 1) for i in range(layers)]
    if layers == 1:
        tex = NeuralTexture.from_tensor(ls[0].clone()).cuda()
    else:
        tex = HierarchicalNeuralTexture.from_tensor([l.clone() for l in ls]).cuda()
    got = export.texture_rgb8(tex).cpu()
    if layers == 1:                        # texture.py:56-57: the single-layer image is the Parameter itself
        x = ls[0].clone().mul_(1.0 / 255)
        x = (x - torch.tensor([-0.40760392, -0.45795686, -0.48501961]).view(3, 1, 1)) / 1.0
        want = x[torch.LongTensor([2, 1, 0])].clamp(0, 1).mul(255).byte().permute(1, 2, 0).contiguous()
    else:
        want = _reference_texture_bytes(ls)
    assert got.shape == want.shape and got.dtype == torch.uint8
    diff = (got.int() - want.int()).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) <= 5e-3, (int(diff.max()), float((diff > 0).float().mean()))
    # the module's save_image(..., post()) takes the device path and writes the same picture as the reference bytes
    # pushed through the same JPEG encoder
    tex.save_image(str(tmp_path), "t_", normalize_transform=post())
    from PIL import Image
    import numpy as np
    Image.fromarray(want.numpy()).save(str(tmp_path / "want.jpg"))
    jpg = np.asarray(Image.open(str(tmp_path / "t_texture.jpg")).convert("RGB")).astype(int)
    ref = np.asarray(Image.open(str(tmp_path / "want.jpg")).convert("RGB")).astype(int)
    assert jpg.shape == ref.shape == tuple(want.shape) and np.abs(jpg - ref).mean() < 0.5, np.abs(jpg - ref).mean()


def test_mip_preview_is_gl_trilinear_of_the_box_filtered_chain():
    """stylemesh_b200.export.MipPreview against torch: mip level l = avg_pool2d^l, GL_LINEAR + CLAMP_TO_EDGE at texel
    centres = grid_sample(align_corners=False, padding_mode='border'), blended by frac(lod); (0, 0) pixels black."""
    import torch.nn.functional as F
    from stylemesh_b200 import export
    from stylemesh_b200.model.texture.texture import HierarchicalNeuralTexture
    g = torch.Generator().manual_seed(3)
    ls = [torch.rand(3, 128 // 2 ** i, 128 // 2 ** i, generator=g) * 200 - 100 for i in range(3)]
    tex = HierarchicalNeuralTexture.from_tensor([l.clone() for l in ls]).cuda()
    mips = export.MipPreview(tex)
    assert [tuple(m.shape[1:]) for m in mips.levels] == [(128 >> i, 128 >> i) for i in range(8)]
    base = mips.levels[0].cpu()
    chain = [base]
    while chain[-1].shape[1] > 1:
        chain.append(F.avg_pool2d(chain[-1].unsqueeze(0), 2)[0])
    for a, b in zip(mips.levels, chain):
        assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-4)
    H, W = 37, 53
    uv = torch.rand(H, W, 3, generator=g)
    uv[..., 2] = uv[..., 2] * 5.5 - 0.5                       # LOD in [-0.5, 5]: clamped below, blended above
    uv[:4, :, :2] = 0.0                                       # no geometry
    got = mips.render(uv.cuda()).cpu()
    grid = (uv[..., :2] * 2 - 1).unsqueeze(0)
    lod = uv[..., 2].clamp(0, len(chain) - 1)
    l0 = lod.floor().long()
    l1 = (l0 + 1).clamp(max=len(chain) - 1)
    t = (lod - l0.float())
    samples = torch.stack([F.grid_sample(c.unsqueeze(0), grid, mode="bilinear", padding_mode="border",
                                         align_corners=False)[0] for c in chain])          # (levels, 3, H, W)
    idx0 = l0.view(1, 1, H, W).expand(1, 3, H, W)
    idx1 = l1.view(1, 1, H, W).expand(1, 3, H, W)
    val = samples.gather(0, idx0)[0] * (1 - t) + samples.gather(0, idx1)[0] * t
    x = val * (1.0 / 255) + torch.tensor([0.40760392, 0.45795686, 0.48501961]).view(3, 1, 1)
    want = (x[[2, 1, 0]].clamp(0, 1) * 255).byte().permute(1, 2, 0)
    want[:4] = 0
    diff = (got.int() - want.int()).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 0.02
    assert int(got[:4].max()) == 0
