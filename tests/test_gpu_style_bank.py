"""Style bank (all affine layers of a synthesis network in one launch, spi_b200/ops/style_bank.py) against the per-layer
FullyConnectedLayer evaluation it replaces (eg3d/training/networks_stylegan2.py:95-133 as called from :316 and :357-358)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _layers(dev, sizes, k=512):
    from spi_b200.training.networks_stylegan2 import FullyConnectedLayer
    torch.manual_seed(3)
    fcs = [FullyConnectedLayer(k, i, bias_init=1).to(dev) for i in sizes]
    for fc in fcs:
        fc.bias.data.add_(0.1 * torch.randn_like(fc.bias))
    return fcs


@pytest.mark.parametrize('n,expand', [(1, False), (4, False), (4, True)])
def test_style_bank_matches_layers(n, expand):
    from spi_b200.ops.style_bank import style_bank
    dev = torch.device('cuda')
    sizes = [512, 512, 512, 256, 256, 96, 40, 32]              # 40: a row count that is not a multiple of the kernel's 32-row chunks
    fcs = _layers(dev, sizes)
    idx = [0, 1, 1, 2, 3, 3, 4, 4]
    gains = [1.0, 1.0, 1 / 512 ** 0.5, 1.0, 1.0, 0.25, 1.0, 2.0]
    ws = torch.randn(1 if expand else n, 5, 512, device=dev, requires_grad=True)
    ws_in = ws.expand(n, -1, -1) if expand else ws             # expand: one latent broadcast over the batch (stride 0)
    outs = style_bank(ws_in, list(zip(fcs, idx, gains)))
    cot = [torch.randn_like(o) for o in outs]
    loss = sum((o * c).sum() for o, c in zip(outs, cot))
    params = [p for fc in fcs for p in (fc.weight, fc.bias)]
    got = torch.autograd.grad(loss, [ws] + params)

    wd = ws.detach().double().requires_grad_(True)
    wsd = wd                                                  # (a broadcast latent yields ONE style row: the layers share one weight set)
    ref_outs = []
    for fc, j, g in zip(fcs, idx, gains):
        ref_outs.append((wsd[:, j] @ (fc.weight.double() * fc.weight_gain).t() + fc.bias.double() * fc.bias_gain) * g)
    for o, r in zip(outs, ref_outs):
        assert o.shape == r.shape
        assert (o.double() - r).abs().max().item() < 2e-5 * max(1.0, r.abs().max().item())
    ref_loss = sum((o * c.double()).sum() for o, c in zip(ref_outs, cot))
    ref = torch.autograd.grad(ref_loss, [wd] + params)
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert (a.double() - b).abs().max().item() < 3e-5 * max(1.0, b.abs().max().item())


def test_style_bank_unused_outputs_and_frozen_latents():
    from spi_b200.ops.style_bank import style_bank
    dev = torch.device('cuda')
    fcs = _layers(dev, [64, 128, 32])
    ws = torch.randn(2, 3, 512, device=dev)                    # no gradient wanted for the latents (PTI stage 2)
    outs = style_bank(ws, [(fcs[0], 0, 1.0), (fcs[1], 1, 1.0), (fcs[2], 2, 1.0)])
    loss = outs[0].square().sum() + outs[2].sum()              # layer 1's styles unused: its gradients are zero
    loss.backward()
    assert fcs[1].weight.grad is None or fcs[1].weight.grad.abs().max().item() == 0
    ref = 2 * outs[0].detach().t() @ ws[:, 0] * fcs[0].weight_gain
    assert (fcs[0].weight.grad - ref).abs().max().item() < 1e-4 * ref.abs().max().item()
    assert (fcs[2].bias.grad - 2.0).abs().max().item() < 1e-6


def test_synthesis_with_bank_equals_per_layer_affines():
    """The network-level switch: SynthesisNetwork.forward with the bank == the same forward with each layer evaluating its own affine."""
    from spi_b200.ops import style_bank as sb
    from spi_b200.training.networks_stylegan2 import SynthesisNetwork
    dev = torch.device('cuda')
    torch.manual_seed(0)
    net = SynthesisNetwork(w_dim=512, img_resolution=32, img_channels=96, channel_base=4096, channel_max=128, num_fp16_res=0).to(dev).eval().requires_grad_(True)
    ws = torch.randn(2, net.num_ws, 512, device=dev, requires_grad=True)
    img = net(ws, noise_mode='const')
    g = torch.autograd.grad(img.square().mean(), [ws] + list(net.parameters()), allow_unused=True)
    usable = sb.usable
    sb.usable = lambda w: False
    try:
        img2 = net(ws, noise_mode='const')
        g2 = torch.autograd.grad(img2.square().mean(), [ws] + list(net.parameters()), allow_unused=True)
    finally:
        sb.usable = usable
    # the styles differ in the last fp32 bit (summation order); the convolutions round their operands to TF32, which turns a last-bit
    # difference of a weight into a 2^-11 one now and then: the images agree to the engine's TF32 tolerance, not to fp32 rounding
    assert ((img - img2).norm() / img2.norm()).item() < 1e-3
    names = ['ws'] + [k for k, _ in net.named_parameters()]
    for name, a, b in zip(names, g, g2):
        if b is None:
            assert a is None or a.abs().max().item() == 0, name
            continue
        if name.endswith('noise_strength'):
            continue        # a scalar sum of signed per-pixel terms that nearly cancel: ill-conditioned (its own tolerance: test_gpu_generator.py)
        # two evaluations of the same network differ by the reduction order of the split / reduce-add convolutions (1e-4 level)
        rel = ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
        assert rel < 1e-2, (name, rel)


def test_modulate_bank_matches_per_layer_modulation():
    """`modulate_bank` (all layers in one launch, transposed copies included) == `modulate_weights` + `tc2_wT` layer by layer: forward
    bit-identical (same kernel body), gradients to fp32 rounding (atomics), including a gradient handed back in the transposed layout the
    stride-2 weight-gradient kernel writes and a layer whose weights go unused."""
    from spi_b200.ops import conv as E
    from spi_b200.ops.modulate import bank_usable, modulate_bank, modulate_weights
    dev = torch.device('cuda')
    gen = torch.Generator().manual_seed(5)
    specs = [  # O, I, k, n, demodulate, layout, flip, transposed
        (512, 512, 3, 1, True, 'ohwi', False, 'rev'), (256, 512, 3, 2, True, 'ohwi', True, 'keep'), (96, 256, 1, 2, False, 'ohwi', False, 'rev'),
        (3, 128, 1, 4, False, 'ohwi', False, None), (64, 32, 3, 1, True, 'oihw', False, None), (128, 128, 3, 1, True, 'ohwi', False, 'rev')]
    weights = [torch.randn(o, i, k, k, generator=gen).to(dev).requires_grad_(True) for o, i, k, *_ in specs]
    styles = [(1 + 0.2 * torch.randn(n, i, generator=gen)).to(dev).requires_grad_(True) for _, i, _, n, *_ in specs]
    entries = [(w, s, d, lay, fl, tr) for w, s, (_, _, _, _, d, lay, fl, tr) in zip(weights, styles, specs)]
    assert bank_usable(entries)
    bank = modulate_bank(entries)
    single = [modulate_weights(w, s, d, layout=lay, flip=fl) for w, s, d, lay, fl, _ in entries]
    cots = []
    for (w5, wT), ref, (o, i, k, n, d, lay, fl, tr) in zip(bank, single, specs):
        assert w5.shape == ref.shape and w5.stride() == ref.stride() and torch.equal(w5, ref)
        if tr is None:
            assert wT is None
        else:
            wk = E._ohwi(ref.detach()).view(n, o, k * k, i)
            assert torch.equal(wT, E.tc2_wT(wk, tr == 'rev'))
        cots.append(torch.randn(ref.shape, generator=gen).to(dev))
    live = [0, 1, 2, 3, 5]                                   # layer 4's weights are not used by the loss
    cot1 = torch.empty(2, 512, 3, 3, 256, device=dev).permute(0, 4, 1, 2, 3).copy_(cots[1])     # layer 1: cotangent laid out [n][i][kh][kw][o]
    loss_b = sum((bank[l][0] * (cot1 if l == 1 else cots[l])).sum() for l in live)
    loss_s = sum((single[l] * cots[l]).sum() for l in live)
    gb = torch.autograd.grad(loss_b, weights + styles, allow_unused=True)
    gs = torch.autograd.grad(loss_s, weights + styles, allow_unused=True)
    for a, b in zip(gb, gs):
        if b is None:
            assert a is None
            continue
        assert ((a - b).norm() / b.norm()).item() < 1e-6
