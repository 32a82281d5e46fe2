"""Camera helpers of the inversion loop (drop-in for the used part of spi/utils/camera_utils.py).

Same functions / signatures / random-draw order as the reference; the per-sample Python loops that call
`math.cos(tensor)` (8 implicit device->host syncs per sampled camera, camera_utils.py:170-193) are replaced by batched
device-side trigonometry, so sampling cameras never synchronises the stream.
"""
import math

import torch

from . import rng


_consts = {}


def _const(values, device):
    """Small constant tensors are uploaded once per device and reused: a host->device copy of pageable memory is not
    allowed while a CUDA graph is being captured."""
    key = (str(values), str(device))
    if key not in _consts:
        _consts[key] = torch.tensor(values, dtype=torch.float32, device=device)
    return _consts[key]


def normalize_vecs(v):
    return v / torch.norm(v, dim=-1, keepdim=True)


def create_cam2world_matrix(forward_vector, origin):
    """camera_utils.py:123-143."""
    forward_vector = normalize_vecs(forward_vector)
    up = _const([0, 1, 0], origin.device).expand_as(forward_vector)
    right = -normalize_vecs(torch.cross(up, forward_vector, dim=-1))
    up = normalize_vecs(torch.cross(forward_vector, right, dim=-1))
    n = forward_vector.shape[0]
    rot = torch.eye(4, device=origin.device).unsqueeze(0).repeat(n, 1, 1)
    rot[:, :3, :3] = torch.stack((right, up, forward_vector), dim=-1)
    trans = torch.eye(4, device=origin.device).unsqueeze(0).repeat(n, 1, 1)
    trans[:, :3, 3] = origin
    return trans @ rot


class LookAtPoseSampler:
    """camera_utils.py:58-89."""

    @staticmethod
    def sample(horizontal_mean, vertical_mean, lookat_position, horizontal_stddev=0, vertical_stddev=0, radius=1, batch_size=1,
               device='cpu', sample_mode='randn'):
        if sample_mode == 'randn':
            h = torch.randn((batch_size, 1), device=device) * horizontal_stddev + horizontal_mean
            v = torch.randn((batch_size, 1), device=device) * vertical_stddev + vertical_mean
        else:   # uniform, one-sided as written in the reference (camera_utils.py:74-76)
            h = rng.rand((batch_size, 1), device) * horizontal_stddev + horizontal_mean
            v = rng.rand((batch_size, 1), device) * vertical_stddev + vertical_mean
        v = torch.clamp(v, 1e-5, math.pi - 1e-5)
        phi = torch.arccos(1 - 2 * (v / math.pi))
        origins = torch.zeros((batch_size, 3), device=device)
        origins[:, 0:1] = radius * torch.sin(phi) * torch.cos(math.pi - h)
        origins[:, 2:3] = radius * torch.sin(phi) * torch.sin(math.pi - h)
        origins[:, 1:2] = radius * torch.cos(phi)
        return create_cam2world_matrix(normalize_vecs(lookat_position - origins), origins)


def _intrinsics(batch_size, device):
    return _const([[4.2647, 0, 0.5], [0, 4.2647, 0.5], [0, 0, 1]], device).view(1, 9).repeat(batch_size, 1)


def sample_camera(batch_size=1, yaw_range=0.35, pitch_range=0.25, device='cpu'):
    """camera_utils.py:159-167."""
    lookat = _const([0, 0, 0.2], device)
    ext = LookAtPoseSampler.sample(horizontal_mean=math.pi / 2, vertical_mean=math.pi / 2 - 0.2, lookat_position=lookat,
                                   horizontal_stddev=yaw_range, vertical_stddev=pitch_range, radius=2.7, batch_size=batch_size,
                                   device=device, sample_mode='uniform')
    return torch.cat([ext.view(-1, 16), _intrinsics(batch_size, device)], dim=1)


def cal_canonical_c(yaw_angle=0, pitch_angle=0, batch_size=1, device='cpu'):
    """camera_utils.py:233-240."""
    lookat = _const([0, 0, 0.2], device)
    ext = LookAtPoseSampler.sample(math.pi / 2 + yaw_angle, math.pi / 2 - 0.2 + pitch_angle, lookat, radius=2.7,
                                   batch_size=batch_size, device=device)
    return torch.cat([ext.view(-1, 16), _intrinsics(batch_size, device)], dim=1)


def angle_to_rotation(yaw, pitch, roll=0):
    """camera_utils.py:170-193, batched on device: R = R_yaw @ R_pitch @ R_roll for [B] tensors of angles."""
    yaw, pitch = yaw.reshape(-1), pitch.reshape(-1)
    roll = torch.zeros_like(yaw) + roll if not torch.is_tensor(roll) else roll.reshape(-1).to(yaw)
    z, o = torch.zeros_like(yaw), torch.ones_like(yaw)
    cy, sy, cp, sp, cr, sr = torch.cos(yaw), torch.sin(yaw), torch.cos(pitch), torch.sin(pitch), torch.cos(roll), torch.sin(roll)
    R_roll = torch.stack([cr, -sr, z, sr, cr, z, z, z, o], -1).view(-1, 3, 3)
    R_yaw = torch.stack([cy, z, sy, z, o, z, -sy, z, cy], -1).view(-1, 3, 3)
    R_pitch = torch.stack([o, z, z, z, cp, -sp, z, sp, cp], -1).view(-1, 3, 3)
    return R_yaw @ R_pitch @ R_roll


def sample_surrounding_camera(middle_camera, batch_size=1, yaw_range=0.1, pitch_range=0.1):
    """camera_utils.py:196-211: the top three rows of cam2world (rotation AND translation) are left-multiplied."""
    device = middle_camera.device
    y = (rng.rand((batch_size, 1), device) * 2 - 1) * yaw_range + 0.0
    p = (rng.rand((batch_size, 1), device) * 2 - 1) * pitch_range + 0.0
    rot = angle_to_rotation(y, p).float()
    cam = middle_camera.repeat(batch_size, 1).clone()
    ext = cam[:, :16].view(-1, 4, 4).clone()
    ext[:, :3] = torch.bmm(rot, ext[:, :3])
    cam[:, :16] = ext.view(-1, 16)
    return cam


def flip_yaw(pose_matrix):
    flipped = pose_matrix.clone()
    for (i, j) in ((0, 1), (0, 2), (0, 3), (1, 0), (2, 0)):
        flipped[:, i, j] *= -1
    return flipped


def cal_mirror_c(camera):
    """camera_utils.py:346-350."""
    pose, intr = camera[:, :16].reshape(-1, 4, 4), camera[:, 16:].reshape(-1, 3, 3)
    return torch.cat([flip_yaw(pose).view(-1, 16), intr.reshape(-1, 9)], dim=1)


def rotation_to_angle(matrix):
    """camera_utils.py:353-364."""
    pitch = torch.arctan(-matrix[1, 2] / matrix[2, 2])
    yaw = torch.arctan(matrix[0, 2] * torch.cos(pitch) / matrix[2, 2])
    roll = torch.arctan(-matrix[0, 1] / matrix[0, 0])
    return yaw, pitch, roll


def gauss_function(x, mean=0.0, std=0.25):
    return torch.exp(-0.5 * (x - mean) * (x - mean) / std / std) / (std * math.sqrt(2 * math.pi))


def cal_camera_gauss_weight(camera):
    """camera_utils.py:388-395."""
    return [gauss_function(rotation_to_angle(c.view(25)[:16].view(4, 4)[:3, :3])[0], std=0.4) / 2.6 for c in camera]


def cal_camera_weight(camera):
    """camera_utils.py:398-412: (1 - N(|yaw|; 0, 0.29)/2.7)/2, zero when |yaw| < 0.2.  Batched, no host branch."""
    weight = []
    for c in camera:
        y = torch.abs(rotation_to_angle(c.view(25)[:16].view(4, 4)[:3, :3])[0])
        w = (1 - gauss_function(y, std=0.29) / 2.7) / 2
        weight.append(torch.where(y < 0.2, torch.zeros_like(w), w))
    return torch.stack(weight, dim=0)
