"""Import shim for the read-only reference tree (`/root/reference`).

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` *in the build container* to run the
reference's own CPU code and mint golden vectors (SURVEY.md §8c recipe).  Nothing on the GPU box
may import this: `/root/reference` does not exist there.

The shim never writes to the reference tree; all patches are applied to live module objects:
  * `Tensor.cuda` / `Module.cuda` become identity and `"cuda"` is stripped from `Module.to`
    (hard-coded device strings at spi/criteria/lpips/lpips.py:25,28, spi/utils/rotate.py:102,108,
    spi/criteria/bbox_cx_loss.py:47-57),
  * torchvision `vgg16` / `vgg19` constructors are forced to `weights=None`
    (spi/criteria/lpips/networks.py:92, spi/criteria/bbox_cx_loss.py:79 download weights),
  * `lpips.utils.get_state_dict` returns caller-supplied lin weights (spi/criteria/lpips/utils.py:13),
  * `imageio` / `mrcfile` are stubbed (spi/utils/video_utils.py:17,22).
"""
import os
import sys
import types

REF = os.environ.get('SPI_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'eg3d'))


def install():
    import torch
    import torchvision
    sys.dont_write_bytecode = True
    for p in (REF, os.path.join(REF, 'eg3d'), os.path.join(REF, 'spi')):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ('imageio', 'mrcfile'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _to = torch.nn.Module.to

    def to(self, *args, **kwargs):
        args = tuple(a for a in args if not (isinstance(a, str) and a.startswith('cuda')))
        if isinstance(kwargs.get('device'), str) and kwargs['device'].startswith('cuda'):
            kwargs.pop('device')
        if not args and not kwargs:
            return self
        return _to(self, *args, **kwargs)
    torch.nn.Module.to = to
    _vgg16, _vgg19 = torchvision.models.vgg16, torchvision.models.vgg19
    torchvision.models.vgg16 = lambda *a, **k: _vgg16(weights=None)
    torchvision.models.vgg19 = lambda *a, **k: _vgg19(weights=None)
    torchvision.models.vgg.vgg19 = lambda *a, **k: _vgg19(weights=None)


FFHQ512_KWARGS = dict(
    z_dim=512, c_dim=25, w_dim=512, img_resolution=512, img_channels=3,
    mapping_kwargs=dict(num_layers=2), sr_num_fp16_res=4,
    sr_kwargs=dict(channel_base=32768, channel_max=512, fused_modconv_default='inference_only'),
    channel_base=32768, channel_max=512, fused_modconv_default='inference_only',
    num_fp16_res=0, conv_clamp=None,
    rendering_kwargs=dict(
        image_resolution=512, disparity_space_sampling=False, clamp_mode='softplus',
        superresolution_module='training.superresolution.SuperresolutionHybrid8XDC',
        c_gen_conditioning_zero=False, c_scale=1.0, superresolution_noise_mode='none',
        density_reg=0.25, density_reg_p_dist=0.004, reg_type='l1', decoder_lr_mul=1.0,
        sr_antialias=True, depth_resolution=48, depth_resolution_importance=48,
        ray_start=2.25, ray_end=3.3, box_warp=1, avg_camera_radius=2.7,
        avg_camera_pivot=[0, 0, 0.2]),
)


def build_reference_generator(depth_resolution=48, depth_resolution_importance=48):
    """eg3d/training/triplane.py:19 with the FFHQ-512 kwargs of SURVEY.md §8c, eval(), nrr=128
    (spi/utils/load_utils.py:25-32)."""
    import copy
    from training.triplane import TriPlaneGenerator
    kw = copy.deepcopy(FFHQ512_KWARGS)
    kw['rendering_kwargs']['depth_resolution'] = depth_resolution
    kw['rendering_kwargs']['depth_resolution_importance'] = depth_resolution_importance
    G = TriPlaneGenerator(**kw).eval().requires_grad_(False)
    G.neural_rendering_resolution = 128
    return G
