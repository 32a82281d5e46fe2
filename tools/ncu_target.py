"""Small eager (no CUDA graphs) run of the bench workload for Nsight Compute: after one warm-up of each step kind, the
NVTX range 'timed' covers 1 mir iteration + 1 RotBbox cycle (i = 0 heavy, 1..3 light).  Not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from spi_b200.configs import global_config

global_config.use_cuda_graphs = False
job = bench.OursJob('cuda:0', bench.synthetic_inputs())
for kind in ('mir', 'rot', 'rot'):
    job.step(kind)
torch.cuda.synchronize()
job.i_rot = 0
torch.cuda.nvtx.range_push('timed')
for kind in ('mir', 'rot', 'rot', 'rot', 'rot'):
    job.step(kind)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
