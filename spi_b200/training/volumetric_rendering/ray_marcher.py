"""Stand-alone ray marcher (drop-in for `training.volumetric_rendering.ray_marcher`, ray_marcher.py:20-62).
Inside `ImportanceRenderer` the marcher is fused into the render kernel; this module exposes the same arithmetic for
callers that composite pre-computed samples (forward only -- no gradient path uses it on the inversion loop)."""
import torch

from ... import _lib


class MipRayMarcher2(torch.nn.Module):
    def run_forward(self, colors, densities, depths, rendering_options):
        assert rendering_options.get('clamp_mode', 'softplus') == 'softplus', "MipRayMarcher only supports `clamp_mode`=`softplus`!"
        if not colors.is_cuda:
            raise RuntimeError('spi_b200.MipRayMarcher2: tensors must reside on a CUDA device')
        n, r, d, c = colors.shape
        colors, densities, depths = colors.detach().float().contiguous(), densities.detach().float().contiguous(), depths.detach().float().contiguous()
        rgb = torch.empty(n, r, c, device=colors.device)
        depth = torch.empty(n, r, 1, device=colors.device)
        weights = torch.empty(n, r, d - 1, 1, device=colors.device)
        minmax = torch.empty(2, dtype=torch.int32, device=colors.device)
        _lib.check(_lib.load().spi_ray_march(_lib.ptr(colors), _lib.ptr(densities), _lib.ptr(depths), n * r, d, c, _lib.ptr(rgb),
                                             _lib.ptr(depth), _lib.ptr(weights), _lib.ptr(minmax), _lib.stream()))
        if rendering_options.get('white_back', False):
            rgb = rgb + 2 * (1 - weights.sum(2))
        return rgb, depth, weights

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)
