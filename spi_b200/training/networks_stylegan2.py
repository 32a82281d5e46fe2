"""StyleGAN2 generator half used as the tri-plane backbone and inside the super-resolution module
(drop-in for `training.networks_stylegan2`, eg3d/training/networks_stylegan2.py:27-552; the discriminator half
:557-795 is not on the inversion path).

Module / parameter / buffer names and constructor signatures are the reference's (state-dict contract, SURVEY.md
§8b.2).  Execution differs:
  * activations stay channels-last fp32 end to end (one texel / pixel = contiguous channels), which is what the
    implicit-GEMM conv engine and the fused renderer consume;
  * modulation + demodulation are folded into per-sample weights (the reference's `fused_modconv` branch, the one that
    actually runs because G is always in eval mode: networks_stylegan2.py:427-428) and each sample is
    a dense conv problem;
  * FIR / bias / activation passes are the sm_100a kernels behind `torch_utils.ops`.
"""
import os

import numpy as np
import torch

from ..ops import conv as conv_engine
from ..ops.modulate import bank_usable, modulate_bank, modulate_weights
from ..ops import style_bank
from ..torch_utils import misc, persistence
from ..torch_utils.ops import bias_act, conv2d_resample, fma, upfirdn2d


# up-sampling layers: FIR + noise + bias + activation in one kernel (SPI_FUSE_BLUR_EPILOGUE=0: the two separate passes, for A/B timing)
FUSE_BLUR_EPILOGUE = os.environ.get('SPI_FUSE_BLUR_EPILOGUE', '1') != '0'


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def _batched_resample_conv(x, w, f, up, padding, flip_weight, epilogue=None, wT=None):
    """conv2d_resample (conv2d_resample.py:48-143) for per-sample weights w [N, O, I, kh, kw]: the two branches a
    generator layer takes (up = 1: plain conv; up = 2: stride-2 transposed conv then 4x4 FIR with gain 4).  `epilogue` (keyword
    arguments of `bias_act.blur_bias_act_noise`) folds the layer's noise / bias / activation pass into the FIR kernel."""
    o, i, kh, kw = w.shape[1:]
    if up == 1:
        return conv_engine.conv2d_per_sample(x, w, padding=[padding, padding], flip_weight=flip_weight, wT=wT)
    fw, fh = upfirdn2d._get_filter_size(f)
    px0 = padding + (fw + up - 1) // 2 - (kw - 1)
    px1 = padding + (fw - up) // 2 - (kw - up)
    py0 = padding + (fh + up - 1) // 2 - (kh - 1)
    py1 = padding + (fh - up) // 2 - (kh - up)
    pxt, pyt = max(min(-px0, -px1), 0), max(min(-py0, -py1), 0)
    y = conv_engine.conv2d_per_sample(x, w, stride=up, padding=[pyt, pxt], transpose=True, flip_weight=(not flip_weight), wT=wT)
    if epilogue is not None:
        return bias_act.blur_bias_act_noise(y, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], fir_gain=up ** 2, **epilogue)
    return upfirdn2d.upfirdn2d(x=y, f=f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2)


def modulation_plan(weight, up, flip_weight):
    """How the fused branch of `modulated_conv2d` lays out the per-sample weights of a layer: (memory layout, taps written reversed, which
    transposed copy the layer's data-gradient convolution reads -- 'rev' / 'keep' on the tc2 engine, else None)."""
    o, i, kh, kw = weight.shape
    tc2 = conv_engine.ENGINE == 'tc2' and i % 32 == 0 and o % 32 == 0 and kh == kw
    if up == 2:
        # the stride-2 transposed convolution consumes [I,O,kh,kw] channels-last weights with the taps reversed when
        # flip_weight is set (conv2d_resample.py:38-40,117): modulate_weights writes them like that directly
        # (cuDNN's transposed convolution wants [I][kh][kw][O]; this library's engine reads [O][kh][kw][I] for every form)
        return ('ohwi' if tc2 else 'ihwo'), bool(flip_weight) and (kh > 1 or kw > 1), ('keep' if tc2 and kh == 3 else None)
    return 'ohwi', (not flip_weight) and (kh > 1 or kw > 1), ('rev' if tc2 and kh in (1, 3) else None)


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True, blur_epilogue=None, conv_epilogue=None, modw=None):
    """networks_stylegan2.py:34-91.  x [N,I,H,W], weight [O,I,kh,kw], styles [N,I].  `blur_epilogue` (up = 2, fused branch only):
    the caller's noise / bias / activation pass, applied inside the FIR kernel that ends the up-sampling convolution.  `conv_epilogue`
    (up = 1, fused branch only; keyword arguments of `conv_engine.conv2d_bias_act`): the same pass applied inside the convolution's
    accumulator read-out.  `modw` (fused branch only): the layer's per-sample weights and their transposed copy, already built by the
    network's modulation bank (`modulate_bank`, one launch for all layers) according to `modulation_plan`."""
    batch_size = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    misc.assert_shape(weight, [out_channels, in_channels, kh, kw])
    misc.assert_shape(x, [batch_size, in_channels, None, None])
    assert styles.ndim == 2 and styles.shape[1] == in_channels and styles.shape[0] in (1, batch_size)     # 1: one style shared by the batch
    if fused_modconv and x.dtype == torch.float32:
        # one launch: w[n,o,i,k] = W*s (*rsqrt(sum (W s)^2 + 1e-8))  (spi_modulate_weights)
        assert down == 1
        kh, kw = weight.shape[2:]
        layout, pre, _ = modulation_plan(weight, up, flip_weight)
        w, wT = modw if modw is not None else (modulate_weights(weight, styles, demodulate, layout=layout, flip=pre), None)
        if up == 2:
            x = _batched_resample_conv(x, w, resample_filter, up, padding, flip_weight and not pre, epilogue=blur_epilogue, wT=wT)
            if blur_epilogue is not None:
                assert noise is None
                return x
        else:
            if conv_epilogue is not None:
                assert noise is None
                return conv_engine.conv2d_bias_act(x, w, wT=wT, **conv_epilogue)
            x = _batched_resample_conv(x, w, resample_filter, up, padding, True, wT=wT)
        if noise is not None:
            x = x + noise
        return x
    w = dcoefs = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(batch_size, 1, -1, 1, 1)           # [N,O,I,kh,kw]
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()                     # [N,O]
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(batch_size, -1, 1, 1, 1)
    if not fused_modconv:                                                            # :70-80 (training-mode branch)
        x = x * styles.to(x.dtype).reshape(batch_size, -1, 1, 1)
        x = conv2d_resample.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down,
                                            padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = fma.fma(x, dcoefs.to(x.dtype).reshape(batch_size, -1, 1, 1), noise.to(x.dtype))
        elif demodulate:
            x = x * dcoefs.to(x.dtype).reshape(batch_size, -1, 1, 1)
        elif noise is not None:
            x = x.add_(noise.to(x.dtype))
        return x
    assert down == 1
    x = _batched_resample_conv(x, w.to(x.dtype), resample_filter, up, padding, flip_weight)
    if noise is not None:
        x = x + noise
    return x


@persistence.persistent_class
class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:     # (x W^T) * gain + b in one GEMM call (gain as alpha)
            return torch.addmm(b.unsqueeze(0), x, self.weight.to(x.dtype).t(), alpha=float(self.weight_gain))
        w = self.weight.to(x.dtype) * self.weight_gain
        return bias_act.bias_act(x.matmul(w.t()), b, act=self.activation)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


@persistence.persistent_class
class MappingNetwork(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None,
                 activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws = z_dim, c_dim, w_dim, num_ws
        self.num_layers, self.w_avg_beta = num_layers, w_avg_beta
        if embed_features is None:
            embed_features = w_dim
        if c_dim == 0:
            embed_features = 0
        if layer_features is None:
            layer_features = w_dim
        features = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for idx in range(num_layers):
            setattr(self, f'fc{idx}', FullyConnectedLayer(features[idx], features[idx + 1], activation=activation,
                                                         lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        x = None
        if self.z_dim > 0:
            misc.assert_shape(z, [None, self.z_dim])
            x = normalize_2nd_moment(z.to(torch.float32))
        if self.c_dim > 0:
            misc.assert_shape(c, [None, self.c_dim])
            y = normalize_2nd_moment(self.embed(c.to(torch.float32)))
            x = torch.cat([x, y], dim=1) if x is not None else y
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if update_emas and self.w_avg_beta is not None:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x

    def extra_repr(self):
        return f'z_dim={self.z_dim:d}, c_dim={self.c_dim:d}, w_dim={self.w_dim:d}, num_ws={self.num_ws:d}'


@persistence.persistent_class
class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True,
                 activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim, self.resolution = in_channels, out_channels, w_dim, resolution
        self.up, self.use_noise, self.activation, self.conv_clamp = up, use_noise, activation, conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def modulation_entry(self, styles):
        """This layer's entry for `modulate_bank`, matching what forward() asks of modulated_conv2d."""
        return (self.weight, styles, True) + modulation_plan(self.weight, self.up, self.up == 1)

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1, styles=None, modw=None):
        assert noise_mode in ['random', 'const', 'none']
        in_resolution = self.resolution // self.up
        misc.assert_shape(x, [None, self.in_channels, in_resolution, in_resolution])
        if styles is None:       # (`styles`: this layer's affine output, already evaluated by the network's style bank)
            if w.shape[0] > 1 and w.stride(0) == 0 and fused_modconv:
                w = w[:1]        # one latent broadcast over the batch: one weight set, one batched convolution
            styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        fuse_noise = self.use_noise and noise_mode == 'const' and x.dtype == torch.float32
        if self.use_noise and noise_mode == 'const' and not fuse_noise:
            noise = self.noise_const * self.noise_strength
        act_gain = self.act_gain * gain
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        if self.up == 2 and fused_modconv and noise is None and x.dtype == torch.float32 and FUSE_BLUR_EPILOGUE:
            # up-sampling layer: transposed conv -> [4x4 FIR + (constant noise) + bias + lrelu*gain + clamp] in one kernel
            epi = dict(b=self.bias.to(x.dtype), act=self.activation, gain=act_gain, clamp=act_clamp)
            if fuse_noise:
                epi.update(noise_const=self.noise_const, noise_strength=self.noise_strength)
            return modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=None, up=self.up, padding=self.padding,
                                    resample_filter=self.resample_filter, flip_weight=False, fused_modconv=True, blur_epilogue=epi, modw=modw)
        if self.up == 1 and fused_modconv and noise is None and x.dtype == torch.float32 and self.padding == self.weight.shape[-1] // 2:
            # non-resampling layer: conv -> (constant noise) + bias + lrelu*gain + clamp in the convolution's own epilogue
            epi = dict(b=self.bias.to(x.dtype), act=self.activation, gain=act_gain, clamp=act_clamp)
            if fuse_noise:
                epi.update(noise_const=self.noise_const, noise_strength=self.noise_strength)
            return modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=None, up=1, padding=self.padding,
                                    resample_filter=self.resample_filter, flip_weight=True, fused_modconv=True, conv_epilogue=epi, modw=modw)
        x = modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=noise, up=self.up, padding=self.padding,
                             resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv,
                             modw=modw if fused_modconv else None)
        if fuse_noise:      # + noise_const*noise_strength + bias -> lrelu*gain -> clamp in one pass
            return bias_act.bias_act_noise(x, self.bias.to(x.dtype), self.noise_const, self.noise_strength, act=self.activation,
                                           gain=act_gain, clamp=act_clamp)
        return bias_act.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=act_gain, clamp=act_clamp)

    def extra_repr(self):
        return (f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, w_dim={self.w_dim:d}, '
                f'resolution={self.resolution:d}, up={self.up}, activation={self.activation:s}')


@persistence.persistent_class
class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim, self.conv_clamp = in_channels, out_channels, w_dim, conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))

    def modulation_entry(self, styles):
        return (self.weight, styles, False) + modulation_plan(self.weight, 1, True)

    def forward(self, x, w, fused_modconv=True, styles=None, modw=None):
        if styles is None:       # (`styles`: affine output times weight_gain, from the network's style bank)
            if w.shape[0] > 1 and w.stride(0) == 0 and fused_modconv:
                w = w[:1]
            styles = self.affine(w) * self.weight_gain
        if fused_modconv and x.dtype == torch.float32:
            epi = dict(b=self.bias.to(x.dtype), act='linear', clamp=self.conv_clamp)
            return modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False, fused_modconv=True, conv_epilogue=epi, modw=modw)
        x = modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False, fused_modconv=fused_modconv)
        return bias_act.bias_act(x, self.bias.to(x.dtype), clamp=self.conv_clamp)

    def extra_repr(self):
        return f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, w_dim={self.w_dim:d}'


def _fused(block, fused_modconv):
    """The fused_modconv decision of _SynthesisBlockBase.forward (networks_stylegan2.py:425-428)."""
    if fused_modconv is None:
        fused_modconv = block.fused_modconv_default
    if fused_modconv == 'inference_only':
        fused_modconv = not block.training
    return bool(fused_modconv)


def plan_blocks(blocks, styles):
    """Per block: (its slice of the precomputed `styles`, its layers' per-sample weights from ONE grouped modulation launch), or (None, None)
    per block when the styles were not precomputed."""
    if styles is None:
        return [(None, None)] * len(blocks)
    per_block, first = [], 0
    for block in blocks:
        count = block.num_conv + block.num_torgb
        per_block.append(styles[first:first + count])
        first += count
    entries = [e for block, s in zip(blocks, per_block) for e in block.modulation_entries(s)]
    modws = modulate_bank(entries) if bank_usable(entries) else None
    plan, first = [], 0
    for s in per_block:
        plan.append((s, modws[first:first + len(s)] if modws is not None else None))
        first += len(s)
    return plan


class _SynthesisBlockBase(torch.nn.Module):
    """Shared body of SynthesisBlock (networks_stylegan2.py:365-464) and SynthesisBlockNoUp (superresolution.py:158-259);
    `block_up` = 2 or 1."""
    block_up = 2

    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=[1, 3, 3, 1], conv_clamp=256, use_fp16=False, fp16_channels_last=False,
                 fused_modconv_default=True, **layer_kwargs):
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        self.in_channels, self.w_dim, self.resolution, self.img_channels = in_channels, w_dim, resolution, img_channels
        self.is_last, self.architecture, self.use_fp16 = is_last, architecture, use_fp16
        self.channels_last = (use_fp16 and fp16_channels_last)
        self.fused_modconv_default = fused_modconv_default
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.num_conv = self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        if in_channels != 0:
            up_kw = dict(up=2, resample_filter=resample_filter) if self.block_up == 2 else {}
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                        channels_last=self.channels_last, **up_kw, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                    channels_last=self.channels_last, **layer_kwargs)
        self.num_conv += 1
        if is_last or architecture == 'skip':
            self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp, channels_last=self.channels_last)
            self.num_torgb += 1
        if in_channels != 0 and architecture == 'resnet':
            raise NotImplementedError("architecture='resnet' is not used by EG3D generators (skip only)")

    def style_entries(self, first_ws):
        """(affine layer, latent index, output gain) of this block's layers in the order forward() consumes its latents."""
        layers = ([self.conv0] if self.in_channels != 0 else []) + [self.conv1]
        entries = [(layer.affine, first_ws + j, 1.0) for j, layer in enumerate(layers)]
        if self.is_last or self.architecture == 'skip':
            entries.append((self.torgb.affine, first_ws + len(layers), self.torgb.weight_gain))
        return entries

    def modulation_entries(self, styles):
        """`modulate_bank` entries of this block's layers for their (precomputed) styles, in the order of `style_entries`."""
        layers = ([self.conv0] if self.in_channels != 0 else []) + [self.conv1]
        if self.is_last or self.architecture == 'skip':
            layers.append(self.torgb)
        return [layer.modulation_entry(s) for layer, s in zip(layers, styles)]

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, update_emas=False, styles=None, modws=None, **layer_kwargs):
        misc.assert_shape(ws, [None, self.num_conv + self.num_torgb, self.w_dim])
        w_iter = iter(ws.unbind(dim=1))
        s_iter = iter(styles) if styles is not None else iter(lambda: None, 0)     # precomputed styles, same order as w_iter
        m_iter = iter(modws) if modws is not None else iter(lambda: None, 0)       # precomputed per-sample weights, likewise
        # Precision policy of this build: fp32 storage, TF32 tensor-core contraction; `use_fp16` only
        # selects the clamp, which the reference also applies in its fp32 fallback (networks_stylegan2.py:421-423).
        if fused_modconv is None:
            fused_modconv = self.fused_modconv_default
        if fused_modconv == 'inference_only':
            fused_modconv = (not self.training)
        in_res = self.resolution // self.block_up
        if self.in_channels == 0:
            x = self.const.to(torch.float32).unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
        else:
            misc.assert_shape(x, [None, self.in_channels, in_res, in_res])
            x = x.to(torch.float32)
        if self.in_channels == 0:
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, styles=next(s_iter), modw=next(m_iter), **layer_kwargs)
        else:
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, styles=next(s_iter), modw=next(m_iter), **layer_kwargs)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, styles=next(s_iter), modw=next(m_iter), **layer_kwargs)
        if img is not None and self.block_up == 2:
            misc.assert_shape(img, [None, self.img_channels, self.resolution // 2, self.resolution // 2])
            img = upfirdn2d.upsample2d(img, self.resample_filter)
        if self.is_last or self.architecture == 'skip':
            y = self.torgb(x, next(w_iter), fused_modconv=fused_modconv, styles=next(s_iter), modw=next(m_iter)).to(torch.float32)
            img = img.add_(y) if img is not None else y
        return x, img

    def extra_repr(self):
        return f'resolution={self.resolution:d}, architecture={self.architecture:s}'


@persistence.persistent_class
class SynthesisBlock(_SynthesisBlockBase):
    block_up = 2


@persistence.persistent_class
class SynthesisNetwork(torch.nn.Module):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=4, **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels, self.num_fp16_res = w_dim, img_resolution, img_channels, num_fp16_res
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        channels = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        fp16_resolution = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for res in self.block_resolutions:
            block = SynthesisBlock(channels[res // 2] if res > 4 else 0, channels[res], w_dim=w_dim, resolution=res,
                                   img_channels=img_channels, is_last=(res == img_resolution), use_fp16=(res >= fp16_resolution),
                                   **block_kwargs)
            self.num_ws += block.num_conv
            if res == img_resolution:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def forward(self, ws, **block_kwargs):
        misc.assert_shape(ws, [None, self.num_ws, self.w_dim])
        ws = ws.to(torch.float32)
        x = img = None
        blocks = [getattr(self, f'b{res}') for res in self.block_resolutions]
        # all affine layers of the network in one launch (ops/style_bank.py) when every block takes the fused (eval-mode) branch
        styles = None
        if style_bank.usable(ws) and all(_fused(b, block_kwargs.get('fused_modconv')) for b in blocks):
            entries, w_idx = [], 0
            for block in blocks:
                entries += block.style_entries(w_idx)
                w_idx += block.num_conv
            styles = style_bank.style_bank(ws, entries)
        plan = plan_blocks(blocks, styles)
        w_idx = 0
        for block, (block_styles, block_modws) in zip(blocks, plan):
            x, img = block(x, img, ws.narrow(1, w_idx, block.num_conv + block.num_torgb), styles=block_styles, modws=block_modws, **block_kwargs)
            w_idx += block.num_conv
        return img

    def extra_repr(self):
        return (f'w_dim={self.w_dim:d}, num_ws={self.num_ws:d}, img_resolution={self.img_resolution:d}, '
                f'img_channels={self.img_channels:d}, num_fp16_res={self.num_fp16_res:d}')


@persistence.persistent_class
class Generator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)
