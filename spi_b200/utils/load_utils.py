"""Generator / feature-extractor loaders (spi/utils/load_utils.py:15-51).

`load_eg3d()` returns a fresh eval-mode, grad-free `TriPlaneGenerator` with `neural_rendering_resolution = 128`
(load_utils.py:25-32).  Sources, chosen by `paths_config.EG3D_PATH`:
  * `*.pt`          a `torch.save({'init_kwargs': ..., 'G': state_dict})` file (what tools/convert_pkl.py writes from the
                    reference pickle in an environment that has the reference's `legacy`/`persistence` loaders);
  * `synthetic[:seed]`  FFHQ-512 architecture with default-initialised weights under `torch.manual_seed(seed)`
                    (no checkpoint exists offline);
  * `*.pkl`         the reference's source-carrying pickle (eg3d/legacy.py:23), read by `spi_b200.legacy.load_network_pkl`
                    without executing the embedded source.
"""
import copy

import torch

from ..configs import global_config, paths_config
from ..training.triplane import TriPlaneGenerator

FFHQ512_KWARGS = dict(
    z_dim=512, c_dim=25, w_dim=512, img_resolution=512, img_channels=3,
    mapping_kwargs=dict(num_layers=2), sr_num_fp16_res=4,
    sr_kwargs=dict(channel_base=32768, channel_max=512, fused_modconv_default='inference_only'),
    channel_base=32768, channel_max=512, fused_modconv_default='inference_only', num_fp16_res=0, conv_clamp=None,
    rendering_kwargs=dict(
        image_resolution=512, disparity_space_sampling=False, clamp_mode='softplus',
        superresolution_module='training.superresolution.SuperresolutionHybrid8XDC', c_gen_conditioning_zero=False,
        c_scale=1.0, superresolution_noise_mode='none', density_reg=0.25, density_reg_p_dist=0.004, reg_type='l1',
        decoder_lr_mul=1.0, sr_antialias=True, depth_resolution=48, depth_resolution_importance=48, ray_start=2.25,
        ray_end=3.3, box_warp=1, avg_camera_radius=2.7, avg_camera_pivot=[0, 0, 0.2]))

DEPTH_OVERRIDE = None   # (depth_resolution, depth_resolution_importance) forced on every generator built here (bench sweeps)
_template = {}      # network_pkl -> (init_kwargs, cpu state dict): unpickle / initialise once, clone per restart_training()


def build_generator(init_kwargs=None, state_dict=None, device=None, seed=None, rendering_kwargs=None):
    """`rendering_kwargs`: the pickled ATTRIBUTE of the source generator; the reference assigns it over the constructor's copy
    after re-instantiating (`G_new.rendering_kwargs = G.rendering_kwargs`, load_utils.py:27-28), so a pickle re-saved with a changed
    depth_resolution / ray_start / box_warp renders with the changed values."""
    kw = copy.deepcopy(init_kwargs or FFHQ512_KWARGS)
    rk = copy.deepcopy(dict(rendering_kwargs)) if rendering_kwargs is not None else None
    for d in (kw['rendering_kwargs'], rk):
        if DEPTH_OVERRIDE is not None and d is not None:
            d['depth_resolution'], d['depth_resolution_importance'] = DEPTH_OVERRIDE
    if seed is not None:
        torch.manual_seed(seed)
    G = TriPlaneGenerator(**kw).eval().requires_grad_(False)
    if state_dict is not None:
        G.load_state_dict(state_dict, strict=True)
    if rk is not None:
        G.rendering_kwargs = rk
    G.neural_rendering_resolution = 128
    return G.to(device or global_config.device)


def load_eg3d(reload_modules=True, device=None, network_pkl=None):
    device = device or global_config.device
    if network_pkl is None:
        network_pkl = paths_config.EG3D_PATH
    if network_pkl not in _template:
        if network_pkl.startswith('synthetic'):
            seed = int(network_pkl.split(':')[1]) if ':' in network_pkl else 0
            G = build_generator(device='cpu', seed=seed)
            _template[network_pkl] = (copy.deepcopy(FFHQ512_KWARGS), {k: v.clone() for k, v in G.state_dict().items()}, None)
        elif network_pkl.endswith('.pt'):
            blob = torch.load(network_pkl, map_location='cpu')
            _template[network_pkl] = (blob.get('init_kwargs', FFHQ512_KWARGS), blob['G'], blob.get('rendering_kwargs'))
        else:
            from .. import legacy                 # source-carrying pickle (eg3d/legacy.py:23) read without executing its source
            with open(network_pkl, 'rb') as f:
                G = legacy.load_network_pkl(f)['G_ema']
            kw = dict(G.init_kwargs)
            assert not G.init_args, 'TriPlaneGenerator pickles carry keyword arguments only'
            _template[network_pkl] = (kw, {k: v.clone() for k, v in G.state_dict().items()}, copy.deepcopy(dict(G.rendering_kwargs)))
    kw, sd, rk = _template[network_pkl]
    return build_generator(kw, sd, device=device, rendering_kwargs=rk)


def load_sg_vgg(device=None):
    """`checkpoints/vgg16.pt` stand-in (load_utils.py:47-51): the restated LPIPS-space VGG16 extractor."""
    from ..criteria.lpips.vgg16_pt import VGG16LPIPSFeatures
    return VGG16LPIPSFeatures().eval().to(device or global_config.device)
