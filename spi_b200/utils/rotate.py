"""Depth-guided 3-D warp (drop-in for `spi.utils.rotate`, spi/utils/rotate.py:92-116).

`rotate(target_camera, target_depth, src_image, src_camera, src_depth, src_mask, EPS)` keeps the reference signature
and returns `(new_rgb [N,3,H,W], depth_mask [N,1,H,W])`; the body is one `spi_rotate` launch
(spi_b200/csrc/warp.cu).  Source tensors that were produced by `.repeat(N, ...)` / `.expand(...)` of a single view are
detected by their data pointer + stride and broadcast inside the kernel instead of being copied.
The reference calls this under `torch.no_grad()` (rot_bbox_cx_coach.py:92); no gradient is defined.
"""
import torch

from .. import _lib


def _batch_view(t, n, inner_shape):
    """Return (tensor, batch_stride_in_elements) for a [n, *inner] or broadcastable [1, *inner] source."""
    t = t.detach().float()
    if t.shape[0] == 1 and n > 1:
        return t.reshape(1, *inner_shape).contiguous(), 0
    t = t.reshape(n, *inner_shape)
    if t.stride(0) == 0:          # expanded view
        return t[:1].contiguous(), 0
    t = t.contiguous()
    return t, t.stride(0)


@torch.no_grad()
def rotate(target_camera, target_depth, src_image, src_camera, src_depth, src_mask=None, EPS=5e-2):
    if not src_image.is_cuda:
        raise RuntimeError('spi_b200.rotate: tensors must reside on a CUDA device (no CPU path in this build)')
    n = target_camera.shape[0]
    res = src_image.shape[-1]
    dres = target_depth.shape[-1]
    tcam = target_camera.detach().float().reshape(n, 25).contiguous()
    tdep = target_depth.detach().float().reshape(n, dres, dres).contiguous()
    scam, scam_bs = _batch_view(src_camera, n, (25,))
    sdep, sdep_bs = _batch_view(src_depth, n, (src_depth.shape[-2], src_depth.shape[-1]))
    assert src_depth.shape[-1] == dres, 'source and target depth maps must share a resolution (128 in the reference)'
    img, img_bs = _batch_view(src_image, n, (3, res, res))
    msk, msk_bs = (None, 0) if src_mask is None else _batch_view(src_mask, n, (res, res))
    rgb = torch.empty(n, 3, res, res, device=img.device)
    mask = torch.empty(n, 1, res, res, device=img.device)
    _lib.check(_lib.load().spi_rotate(_lib.ptr(tcam), _lib.ptr(tdep), _lib.ptr(img), _lib.ptr(scam), _lib.ptr(sdep), _lib.ptr(msk),
                                      _lib.ptr(rgb), _lib.ptr(mask), n, res, dres, scam_bs, sdep_bs, img_bs, msk_bs, float(EPS),
                                      _lib.stream()))
    return rgb, mask
