"""Oracle (CPU, test infrastructure): cameras, masks and the depth-guided 3-D warp.

Follows:
  look-at pose / cam2world    spi/utils/camera_utils.py:58-89,123-143
  sample_camera               spi/utils/camera_utils.py:159-167
  angle_to_rotation           spi/utils/camera_utils.py:170-193
  sample_surrounding_camera   spi/utils/camera_utils.py:196-211
  mirror camera               spi/utils/camera_utils.py:336-350
  yaw -> mirror weight        spi/utils/camera_utils.py:353-364,383-412
  face mask                   spi/utils/mask_utils.py:4-9
  unproject/project/rotate    spi/utils/rotate.py:5-116
Random draws are injected (`rand=`) so parity does not depend on generator state.
"""
import math

import torch
import torch.nn.functional as F


def _unit(v):
    return v / torch.norm(v, dim=-1, keepdim=True)


def cam2world_from_forward(forward, origin):
    """create_cam2world_matrix (camera_utils.py:123-143): y-up, no roll."""
    forward = _unit(forward)
    up = torch.tensor([0., 1., 0.]).expand_as(forward)
    right = -_unit(torch.cross(up, forward, dim=-1))
    up = _unit(torch.cross(forward, right, dim=-1))
    n = forward.shape[0]
    rot = torch.eye(4).repeat(n, 1, 1)
    rot[:, :3, :3] = torch.stack((right, up, forward), dim=-1)
    trans = torch.eye(4).repeat(n, 1, 1)
    trans[:, :3, 3] = origin
    return trans @ rot


def look_at_pose(h, v, lookat, radius):
    """LookAtPoseSampler.sample after the random draw (camera_utils.py:78-89): h,v [B,1]."""
    v = torch.clamp(v, 1e-5, math.pi - 1e-5)
    phi = torch.arccos(1 - 2 * (v / math.pi))
    origin = torch.zeros(h.shape[0], 3)
    origin[:, 0:1] = radius * torch.sin(phi) * torch.cos(math.pi - h)
    origin[:, 2:3] = radius * torch.sin(phi) * torch.sin(math.pi - h)
    origin[:, 1:2] = radius * torch.cos(phi)
    return cam2world_from_forward(_unit(lookat - origin), origin)


INTRINSICS = (4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1)


def sample_camera(rand, yaw_range=0.35, pitch_range=0.25):
    """camera_utils.py:159-167 with 'uniform' mode (:74-76): h = rand*std + mean (one-sided, as written).
    rand [B,2] in [0,1): column 0 -> h, column 1 -> v."""
    b = rand.shape[0]
    h = rand[:, 0:1] * yaw_range + math.pi / 2
    v = rand[:, 1:2] * pitch_range + (math.pi / 2 - 0.2)
    ext = look_at_pose(h, v, torch.tensor([0., 0., 0.2]), 2.7)
    intr = torch.tensor([INTRINSICS]).repeat(b, 1)
    return torch.cat([ext.reshape(-1, 16), intr], 1)


def yaw_pitch_rotation(yaw, pitch):
    """angle_to_rotation with roll=0 (camera_utils.py:170-193): R = R_yaw(y-axis) @ R_pitch(x-axis); float64
    matrices cast to float32 afterwards (:204)."""
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    rp = torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=torch.float64)
    return ry @ rp


def sample_surrounding_camera(camera, rand, yaw_range=0.1, pitch_range=0.1):
    """camera_utils.py:196-211: left-multiply the top three rows of cam2world (rotation AND translation) by
    R(yaw, pitch).  rand [B,2]: column 0 -> yaw draw, column 1 -> pitch draw."""
    b = rand.shape[0]
    y = (rand[:, 0] * 2 - 1) * yaw_range
    p = (rand[:, 1] * 2 - 1) * pitch_range
    rot = torch.stack([yaw_pitch_rotation(float(a), float(c)) for a, c in zip(y, p)]).float()
    cam = camera.repeat(b, 1).clone()
    ext = cam[:, :16].reshape(-1, 4, 4).clone()
    ext[:, :3] = torch.bmm(rot, ext[:, :3])
    cam[:, :16] = ext.reshape(-1, 16)
    return cam


def mirror_camera(camera):
    """cal_mirror_c / flip_yaw (camera_utils.py:336-350)."""
    pose = camera[:, :16].reshape(-1, 4, 4).clone()
    for (i, j) in ((0, 1), (0, 2), (0, 3), (1, 0), (2, 0)):
        pose[:, i, j] *= -1
    return torch.cat([pose.reshape(-1, 16), camera[:, 16:].reshape(-1, 9)], 1)


def camera_yaw(camera_row):
    """rotation_to_angle yaw (camera_utils.py:353-364)."""
    r = camera_row.reshape(25)[:16].reshape(4, 4)[:3, :3]
    pitch = torch.arctan(-r[1, 2] / r[2, 2])
    return torch.arctan(r[0, 2] * torch.cos(pitch) / r[2, 2])


def camera_weight(camera):
    """cal_camera_weight (camera_utils.py:398-412): (1 - N(|yaw|; 0, 0.29)/2.7)/2, zero when |yaw| < 0.2."""
    out = []
    for c in camera:
        y = torch.abs(camera_yaw(c))
        g = torch.exp(-0.5 * y * y / 0.29 / 0.29) / (0.29 * math.sqrt(2 * math.pi))
        w = (1 - g / 2.7) / 2
        out.append(torch.zeros_like(w) if y < 0.2 else w)
    return torch.stack(out)


def face_mask(parsing):
    """calculate_face_mask (mask_utils.py:4-9)."""
    out = torch.zeros_like(parsing)
    for att in (1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13):
        out += (parsing == att)
    return out


# ----------------------------------------------------------------------------- depth-guided warp

def _pixel_grid(n, res):
    t = torch.arange(res, dtype=torch.float32) * (1. / res) + (0.5 / res)
    y = t[:, None].expand(res, res).reshape(1, -1).expand(n, -1)
    x = t[None, :].expand(res, res).reshape(1, -1).expand(n, -1)
    return x, y


def unproject(depth, cam2world, intr, res):
    """rotate.py:5-29: pixel centres + depth -> homogeneous world points [N,M,4]."""
    n = cam2world.shape[0]
    fx, fy = intr[:, 0, 0:1], intr[:, 1, 1:2]
    cx, cy, sk = intr[:, 0, 2:3], intr[:, 1, 2:3], intr[:, 0, 1:2]
    x, y = _pixel_grid(n, res)
    z = depth.reshape(n, res * res)
    xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    yl = (y - cy) / fy * z
    pts = torch.stack((xl, yl, z, torch.ones_like(z)), -1)
    return torch.bmm(cam2world, pts.permute(0, 2, 1)).permute(0, 2, 1)


def project(world, cam2world, intr):
    """rotate.py:32-52: world -> (uv in [0,1], camera-space z)."""
    fx, fy = intr[:, 0, 0:1], intr[:, 1, 1:2]
    cx, cy, sk = intr[:, 0, 2:3], intr[:, 1, 2:3], intr[:, 0, 1:2]
    cam = torch.bmm(torch.inverse(cam2world), world.permute(0, 2, 1)).permute(0, 2, 1)
    xl, yl, z = cam[:, :, 0], cam[:, :, 1], cam[:, :, 2]
    v = (yl / z * fy) + cy
    u = xl / z * fx + sk * v / fy - cy * sk / fy + cx
    return torch.stack([u, v], -1), z


def rotate(target_camera, target_depth, src_image, src_camera, src_depth, src_mask=None, eps=5e-2):
    """rotate / __rotate (rotate.py:56-116) -> (warped rgb [N,3,H,W], mask [N,1,H,W])."""
    n = src_image.shape[0]
    res = src_image.shape[-1]
    tex, tin = target_camera[:, :16].reshape(n, 4, 4), target_camera[:, 16:].reshape(n, 3, 3)
    gex, gin = src_camera[:, :16].reshape(n, 4, 4), src_camera[:, 16:].reshape(n, 3, 3)

    def up(d):
        d = d.reshape(n, 1, 128, 128)
        if res != 128:
            d = F.interpolate(d, (res, res), mode='bilinear', align_corners=False)
        return d.reshape(n, res, res)
    td, gd = up(target_depth), up(src_depth)
    world = unproject(td, tex, tin, res)
    uv, z = project(world, gex, gin)
    grid = 2 * uv.reshape(n, res, res, 2) - 1
    inb = 1 - ((grid[..., 0] < -1) | (grid[..., 0] > 1) | (grid[..., 1] < -1) | (grid[..., 1] > 1)).float()
    src_d = F.grid_sample(gd.reshape(n, 1, res, res), grid, align_corners=False).reshape(n, res, res)
    dmask = ((torch.abs(src_d - z.reshape(n, res, res)) < eps) * inb).unsqueeze(1)
    rgb = F.grid_sample(src_image, grid, align_corners=False) * dmask
    if src_mask is not None:
        m = F.grid_sample(src_mask.reshape(n, 1, res, res), grid, align_corners=False)
        rgb = rgb * m
        dmask = dmask * m
    return rgb, dmask
