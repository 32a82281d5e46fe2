"""Ray generation (drop-in for `training.volumetric_rendering.ray_sampler`, ray_sampler.py:18-61)."""
import torch

from ... import _lib


class RaySampler(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ray_origins_h, self.ray_directions, self.depths, self.image_coords, self.rendering_options = None, None, None, None, None

    def forward(self, cam2world_matrix, intrinsics, resolution):
        """cam2world [N,4,4], intrinsics [N,3,3] -> ray_origins [N,M,3], ray_dirs [N,M,3]; ray m = row*res + col.
        Cameras are constants on the inversion path, so no gradient is produced."""
        if not cam2world_matrix.is_cuda:
            raise RuntimeError('spi_b200.RaySampler: cameras must reside on a CUDA device (no CPU path in this build)')
        n = cam2world_matrix.shape[0]
        cam = torch.cat([cam2world_matrix.reshape(n, 16), intrinsics.reshape(n, 9)], 1).detach().float().contiguous()
        origins = torch.empty(n, resolution * resolution, 3, device=cam.device)
        dirs = torch.empty_like(origins)
        _lib.check(_lib.load().spi_ray_sampler(_lib.ptr(cam), n, resolution, _lib.ptr(origins), _lib.ptr(dirs), _lib.stream()))
        return origins, dirs
