"""CLI / config surface of the inversion (drop-in for spi/run_inversion.py:16-129): the same 21 flags, copied into the
module-global `hyperparameters` / `paths_config`, and the same coach dispatch on --G_1_type.

Extra, B200-specific: when launched under torchrun (RANK/WORLD_SIZE set) each rank takes the contiguous
`--dataset_block (rank+1)/world_size` slice of the images and binds cuda:LOCAL_RANK -- the images are independent, so
there is no data-path collective."""
import argparse
import os

import torch
from torch.utils.data import DataLoader

from .configs import global_config, hyperparameters, paths_config


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Training')
    parser.add_argument('--data_root', type=str, default='test/dataset/')
    parser.add_argument('--data_mode', type=str, default='png')
    parser.add_argument('--output_root', type=str, default=None)
    parser.add_argument('--use_encoder', action='store_true', default=False)
    parser.add_argument('--use_G_avg', action='store_true', default=False)
    parser.add_argument('--use_adapt_yaw_range', action='store_true', default=False)
    parser.add_argument('--not_use_wandb', action='store_true', default=False)
    parser.add_argument('--first_inv_type', type=str, default='pti')
    parser.add_argument('--first_inv_steps', type=int, default=500)
    parser.add_argument('--G_1_step', type=int, default=500)
    parser.add_argument('--G_1_type', type=str, default='space')
    parser.add_argument('--G_2_step', type=int, default=500)
    parser.add_argument('--load_embedding_coach_name', type=str, default=None)
    parser.add_argument('--pt_rot_lambda', type=float, default=0)
    parser.add_argument('--pt_mirror_rot_lambda', type=float, default=0)
    parser.add_argument('--pt_depth_lambda', type=float, default=0)
    parser.add_argument('--pt_tv_lambda', type=float, default=0)
    parser.add_argument('--description', type=str, default=None)
    parser.add_argument('--dataset_block', type=str, default=None, help='1/20')
    parser.add_argument('--select_range', type=int, default=None, help='100')
    parser.add_argument('--filter_index', type=str, default=None, help='1,2,3')
    parser.add_argument('--network', type=str, default=None, help="generator source (spi_b200 extra): *.pt or 'synthetic[:seed]'")
    args = parser.parse_args(argv)
    for k in ('use_encoder', 'use_G_avg', 'first_inv_type', 'first_inv_steps', 'G_1_step', 'G_1_type', 'G_2_step',
              'load_embedding_coach_name', 'use_adapt_yaw_range', 'description', 'pt_rot_lambda', 'pt_mirror_rot_lambda',
              'pt_depth_lambda', 'pt_tv_lambda'):
        setattr(hyperparameters, k, getattr(args, k))
    if args.network is not None:
        paths_config.EG3D_PATH = args.network
    if args.output_root is not None:
        paths_config.root = args.output_root
        paths_config.checkpoints_dir = paths_config.root + 'checkpoints/'
        paths_config.embedding_base_dir = paths_config.root + 'embedding/'
        paths_config.experiments_output_dir = paths_config.root + 'experiments/'
        paths_config.images_output_dir = paths_config.root + 'image/'
        paths_config.mirror_images_output_dir = paths_config.root + 'image_m/'
        paths_config.video_output_dir = paths_config.root + 'video/'
        for d in (paths_config.checkpoints_dir, paths_config.embedding_base_dir, paths_config.experiments_output_dir,
                  paths_config.images_output_dir, paths_config.mirror_images_output_dir, paths_config.video_output_dir):
            os.makedirs(d, exist_ok=True)
    return args


def shard_for_rank(args):
    """torchrun launch: rank r of W -> `--dataset_block (r+1)/W` (images_dataset.py:149-158) on cuda:LOCAL_RANK."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and args.dataset_block is None:
        args.dataset_block = f"{int(os.environ['RANK']) + 1}/{world}"
        global_config.device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
        if torch.cuda.is_available():
            torch.cuda.set_device(global_config.device)
    return args


def build_dataset(args):
    from .data.images_dataset import PTIDataset
    root = args.data_root
    if args.filter_index is not None:
        args.filter_index = args.filter_index.split(',')
    dataset = PTIDataset(source_root=os.path.join(root, 'crop'), c_root=os.path.join(root, 'c'), w_root=None,
                         mask_root=os.path.join(root, 'mask'), lm_root=os.path.join(root, 'lm'), target_name='target',
                         mode=args.data_mode, dataset_block=args.dataset_block, select_range=args.select_range,
                         filter_index=args.filter_index)
    return dataset, DataLoader(dataset, batch_size=1, shuffle=False)


def run(argv=None):
    args = shard_for_rank(parse_args(argv))
    use_wandb = not args.not_use_wandb
    _, dataloader = build_dataset(args)
    from .training.coaches.inference_coach import InferenceCoach
    from .training.coaches.pti_coach import SingleIDCoach
    from .training.coaches.rot_bbox_cx_coach import RotBboxCoach
    if args.G_1_type == 'pti':
        coach = SingleIDCoach(dataloader, use_wandb)
    elif args.G_1_type == 'RotBbox':
        coach = RotBboxCoach(dataloader, use_wandb)
    elif args.G_1_type == 'Inference':
        coach = InferenceCoach(dataloader, use_wandb)
    else:
        raise NotImplementedError
    coach.train()
    return global_config.run_name


if __name__ == '__main__':
    run()
