"""Reproducer: config 2 (sg -> pti) eagerly, synchronising after every step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from spi_b200.configs import global_config
from spi_b200 import _lib
global_config.use_cuda_graphs = '--graphs' in sys.argv
job = bench.OursJob('cuda:0', bench.synthetic_inputs(), 'sg', 'pti', (32, 32), 128)
for i, item in enumerate([('mir', 0), ('mir', 1), ('mir', 2), ('rot', 0), ('rot', 1), ('rot', 2), ('mir', 3), ('rot', 3)]):
    r = job.step(item)
    torch.cuda.synchronize()
    print(i, item, float(r if not isinstance(r, dict) else 0), 'tc err', _lib.load().spi_tc_error(), flush=True)
print('done')
