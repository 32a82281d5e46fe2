"""`PTIDataset` (spi/data/images_dataset.py:102-198): crop/<name>/target.<mode>, c/<name>/target.npy (25 f32),
mask/<name>/target.pt (int64 [1,1,512,512] parsing labels), lm/<name>/target.npy (68x2 at 256 px); `dataset_block i/n`
contiguous sharding (:149-158) is the multi-GPU hook (one block per rank)."""
import glob
import os

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset


def _to_tensor_normalised(img):
    a = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
    return (a - 0.5) / 0.5


class PTIDataset(Dataset):
    def __init__(self, source_root, source_transform=None, c_root=None, w_root=None, mask_root=None, lm_root=None,
                 target_name='target', mode='jpg', dataset_block=None, output_root=None, select_range=None, filter_index=None):
        self.source_root, self.c_root, self.mask_root, self.w_root, self.lm_root = source_root, c_root, mask_root, w_root, lm_root
        self.source_transform = source_transform or _to_tensor_normalised
        self.mode, self.target_name = mode, target_name
        self.source_paths = sorted(glob.glob(f'{source_root}/*/'))
        if select_range is not None:
            self.source_paths = self.source_paths[:select_range]
        if output_root is not None:
            done = set(sorted(glob.glob(f'{output_root}/*.jpg')))
            self.source_paths = [p for p in self.source_paths if os.path.join(output_root, f"{p.split('/')[-2]}.jpg") not in done]
        if dataset_block is not None:
            index, total = (int(v) for v in dataset_block.split('/'))
            block = len(self.source_paths) // total + 1
            self.source_paths = self.source_paths[(index - 1) * block: index * block]
        if filter_index is not None:
            self.source_paths = [os.path.join(source_root, f'{ff}/') for ff in filter_index]

    def __len__(self):
        return len(self.source_paths)

    def __getitem__(self, index):
        path = self.source_paths[index]
        name = os.path.dirname(path).split('/')[-1]
        img = Image.open(os.path.join(path, f'{self.target_name}.{self.mode}')).convert('RGB').resize((512, 512))
        data = {'img': self.source_transform(img), 'fname': self.target_name, 'name': name,
                'c': np.load(os.path.join(self.c_root, name, self.target_name + '.npy')).astype(np.float32)}
        if self.w_root is not None:
            data['w'] = torch.load(os.path.join(self.w_root, name, self.target_name + '.pt'))
        if self.mask_root is not None:
            data['mask'] = torch.load(os.path.join(self.mask_root, name, self.target_name + '.pt'))
        if self.lm_root is not None:
            data['lm'] = torch.from_numpy(np.load(os.path.join(self.lm_root, name, self.target_name + '.npy'))).float()
        return data
