"""Golden vectors for the post-processing rows next to the hot path: the orbit-video camera
trajectory, the frame layout and the `metric_log.txt` text.

TEST INFRASTRUCTURE ONLY; runs in the build container, where `/root/reference` exists:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_post

It executes the reference's own code (spi/utils/video_utils.py:153-160 pose expressions with eg3d/camera_utils.py:58-86,
video_utils.py:30-45 `layout_grid`, spi/training/coaches/base_coach.py:156-198 `log_metric`) and writes
`tests/golden/post.npz` + `tests/golden/metric_log.json`.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    assert ref_shim.available(), 'the reference tree is needed to mint goldens'
    ref_shim.install()
    from eg3d.camera_utils import LookAtPoseSampler
    num = 120
    lookat = torch.tensor([0, 0, 0.2])
    poses = []
    for frame_idx in range(num):          # the expressions of video_utils.py:155-157, num_keyframes = 1, w_frames = 120
        pitch_range, yaw_range = 0.4, 0.7
        p = LookAtPoseSampler.sample(3.14 / 2 + yaw_range * np.sin(2 * 3.14 * frame_idx / (1 * num)),
                                     3.14 / 2 - 0.05 + pitch_range * np.cos(2 * 3.14 * frame_idx / (1 * num)),
                                     lookat, radius=2.7, device='cpu')
        poses.append(p.squeeze().numpy())
    poses = np.stack(poses).astype(np.float32)

    from spi.utils.video_utils import layout_grid
    g = torch.Generator().manual_seed(11)
    imgs = torch.rand(6, 3, 5, 4, generator=g) * 2.4 - 1.2
    grid = layout_grid(imgs, grid_w=3, grid_h=2)
    np.savez_compressed(os.path.join(OUT, 'post.npz'), orbit_poses=poses, grid_in=imgs.numpy(), grid_out=grid)

    from spi.configs import hyperparameters, paths_config
    from spi.training.coaches.base_coach import BaseCoach
    rs = np.random.RandomState(3)
    keys = ('l2', 'lpips', 'id', 'l2_m', 'lpips_m', 'id_m')
    metric_dic = {mode: {k: [float(v) for v in rs.rand(n)] for k in keys} for mode, n in (('w_inv', 5), ('G1_inv', 5))}
    hp = dict(use_encoder=False, first_inv_type='mir', first_inv_steps=500, G_1_step=1000, G_2_step=500)
    for k, v in hp.items():
        setattr(hyperparameters, k, v)
    with tempfile.TemporaryDirectory() as d:
        paths_config.experiments_output_dir = d
        fake = types.SimpleNamespace(coach_name='RotBboxCoach_mir_500_RotBbox_1000', metric_dic=metric_dic)
        BaseCoach.log_metric(fake)
        text = open(os.path.join(d, 'metric_log.txt')).read()
    with open(os.path.join(OUT, 'metric_log.json'), 'w') as fh:
        json.dump({'coach_name': fake.coach_name, 'hyperparameters': hp, 'metric_dic': metric_dic, 'text': text}, fh, indent=1)
    # PTIDataset items and slicing rules, read by the reference's own class from a toy tree; coach names from build_name
    from spi.data.images_dataset import PTIDataset
    from . import toy_dataset
    ds_gold = {}
    with tempfile.TemporaryDirectory() as d:
        toy_dataset.write(d, n=7)
        kw = dict(source_root=os.path.join(d, 'crop'), c_root=os.path.join(d, 'c'), mask_root=os.path.join(d, 'mask'),
                  lm_root=os.path.join(d, 'lm'), mode='png')
        ds = PTIDataset(**kw)
        ds_gold['items'] = [toy_dataset.digest(ds[i]) for i in (0, 6)]
        names = lambda ds: [os.path.dirname(p).split('/')[-1] for p in ds.source_paths]
        ds_gold['blocks'] = {b: names(PTIDataset(dataset_block=b, **kw)) for b in ('1/2', '2/2', '1/3', '3/3', '4/4')}
        ds_gold['select_3'] = names(PTIDataset(select_range=3, **kw))
        ds_gold['filter'] = names(PTIDataset(filter_index=['00004', '00001'], **kw))
        out = os.path.join(d, 'done')
        os.makedirs(out)
        for nm in ('00000', '00003'):
            open(os.path.join(out, nm + '.jpg'), 'w').close()
        ds_gold['resume'] = names(PTIDataset(output_root=out, **kw))
        ds_gold['resume_block_2_2'] = names(PTIDataset(output_root=out, dataset_block='2/2', **kw))
    coach_names = []
    variants = [dict(), dict(use_encoder=True, use_G_avg=True), dict(use_adapt_yaw_range=True, description='abl'),
                dict(first_inv_type='sg', G_1_type='pti', pt_rot_lambda=0, pt_mirror_rot_lambda=0, pt_depth_lambda=0, pt_tv_lambda=0.5)]
    base = dict(first_inv_type='mir', first_inv_steps=500, G_1_type='RotBbox', G_1_step=1000, use_encoder=False, use_G_avg=False,
                pt_rot_lambda=0.1, pt_mirror_rot_lambda=0.05, pt_depth_lambda=1.0, pt_tv_lambda=0.0, use_adapt_yaw_range=False,
                description=None)
    with tempfile.TemporaryDirectory() as d:
        for k in ('checkpoints_dir', 'embedding_base_dir', 'experiments_output_dir', 'images_output_dir', 'mirror_images_output_dir',
                  'video_output_dir'):
            setattr(paths_config, k, os.path.join(d, k) + '/')
        for v in variants:
            cfg = dict(base, **v)
            for k, val in cfg.items():
                setattr(hyperparameters, k, val)
            fake = types.SimpleNamespace(coach_name='RotBboxCoach')
            BaseCoach.build_name(fake)
            coach_names.append({'hyperparameters': cfg, 'coach_name': fake.coach_name,
                                'dirs': sorted(os.path.relpath(os.path.join(r, x), d) for r, dd, _ in os.walk(d) for x in dd)})
            for r, dd, _ in os.walk(d, topdown=False):
                for x in dd:
                    os.rmdir(os.path.join(r, x))
    with open(os.path.join(OUT, 'host_logic.json'), 'w') as fh:
        json.dump({'dataset': ds_gold, 'coach_names': coach_names}, fh)
    print('orbit poses', poses.shape, 'grid', grid.shape, 'metric_log', len(text), 'bytes')


if __name__ == '__main__':
    sys.exit(main())
