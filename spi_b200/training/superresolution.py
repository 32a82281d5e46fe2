"""Super-resolution module of the FFHQ-512 generator (drop-in for `training.superresolution`,
eg3d/training/superresolution.py:158-290): two StyleGAN2 synthesis blocks 32->256 @256^2 and 256->128 @512^2 driven by
the last w, skip-connected RGB."""
import torch

from ..torch_utils import persistence
from ..ops import style_bank
from .networks_stylegan2 import SynthesisBlock, _SynthesisBlockBase, _fused, plan_blocks


@persistence.persistent_class
class SynthesisBlockNoUp(_SynthesisBlockBase):
    """superresolution.py:158-259: same block without the 2x up-sampling of conv0 / the skip image."""
    block_up = 1


@persistence.persistent_class
class SuperresolutionHybrid8XDC(torch.nn.Module):
    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None,
                 channel_base=None, channel_max=None, **block_kwargs):
        super().__init__()
        assert img_resolution == 512
        use_fp16 = sr_num_fp16_res > 0
        self.input_resolution = 128
        self.sr_antialias = sr_antialias
        self.block0 = SynthesisBlock(channels, 256, w_dim=512, resolution=256, img_channels=3, is_last=False,
                                     use_fp16=use_fp16, conv_clamp=(256 if use_fp16 else None), **block_kwargs)
        self.block1 = SynthesisBlock(256, 128, w_dim=512, resolution=512, img_channels=3, is_last=True,
                                     use_fp16=use_fp16, conv_clamp=(256 if use_fp16 else None), **block_kwargs)

    def forward(self, rgb, x, ws, **block_kwargs):
        ws = ws[:, -1:, :].expand(-1, 3, -1)        # a view: keeps a broadcast batch (stride 0) recognisable downstream
        if x.shape[-1] != self.input_resolution:
            size = (self.input_resolution, self.input_resolution)
            x = torch.nn.functional.interpolate(x, size=size, mode='bilinear', align_corners=False, antialias=self.sr_antialias)
            rgb = torch.nn.functional.interpolate(rgb, size=size, mode='bilinear', align_corners=False, antialias=self.sr_antialias)
        styles = None
        if style_bank.usable(ws) and all(_fused(b, block_kwargs.get('fused_modconv')) for b in (self.block0, self.block1)):
            # the six affine layers of the two blocks in one launch (both blocks read the same three identical latents), then their six
            # weight modulations in one more
            styles = style_bank.style_bank(ws, self.block0.style_entries(0) + self.block1.style_entries(0))
        (styles0, modws0), (styles1, modws1) = plan_blocks([self.block0, self.block1], styles)
        x, rgb = self.block0(x, rgb, ws, styles=styles0, modws=modws0, **block_kwargs)
        x, rgb = self.block1(x, rgb, ws, styles=styles1, modws=modws1, **block_kwargs)
        return rgb
