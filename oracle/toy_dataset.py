"""A tiny on-disk dataset in the layout `PTIDataset` reads (spi/data/images_dataset.py:102-198): crop/<name>/target.<mode>,
c/<name>/target.npy, mask/<name>/target.pt, lm/<name>/target.npy.  TEST INFRASTRUCTURE ONLY: shared by
`oracle/make_golden_post.py` (which reads it through the reference's dataset class) and `tests/test_host_logic.py`."""
import os

import numpy as np
import torch
from PIL import Image


def write(root, n=5, mode='png', seed=9):
    rs = np.random.RandomState(seed)
    names = [f'{i:05d}' for i in range(n)]
    for k, name in enumerate(names):
        for sub in ('crop', 'c', 'mask', 'lm'):
            os.makedirs(os.path.join(root, sub, name), exist_ok=True)
        img = np.kron(rs.randint(0, 256, (9, 7, 3)), np.ones((5, 6, 1))).astype(np.uint8)      # 45 x 42: gets resized to 512^2
        Image.fromarray(img).save(os.path.join(root, 'crop', name, f'target.{mode}'))
        np.save(os.path.join(root, 'c', name, 'target.npy'), rs.randn(25))                          # float64 on disk
        torch.save(torch.from_numpy(rs.randint(0, 19, (1, 512, 512))), os.path.join(root, 'mask', name, 'target.pt'))
        np.save(os.path.join(root, 'lm', name, 'target.npy'), rs.rand(68, 2) * 256)
    return names


def digest(item):
    """Small, exact fingerprint of one dataset item."""
    img = item['img']
    return {'img_shape': list(img.shape), 'img_sum': float(img.double().sum()), 'img_grid': img[:, ::64, ::64].numpy().tolist(),
            'c': np.asarray(item['c']).tolist(), 'c_dtype': str(np.asarray(item['c']).dtype),
            'mask_sum': int(item['mask'].sum()), 'mask_dtype': str(item['mask'].dtype), 'mask_shape': list(item['mask'].shape),
            'lm': item['lm'].numpy().tolist(), 'lm_dtype': str(item['lm'].dtype), 'name': item['name'], 'fname': item['fname']}
