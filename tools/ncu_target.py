"""Small eager (no CUDA graphs) run of the bench workload for Nsight Compute: after a warm-up of each step kind (which also
absorbs one-time allocations), cudaProfilerStart/Stop bracket 1 mir iteration + 1 RotBbox cycle (i = 0 heavy, 1..3 light);
run ncu with `--profile-from-start off` so that only those launches (all threads, incl. autograd's) are listed.  Not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from spi_b200.configs import global_config

global_config.use_cuda_graphs = False
job = bench.OursJob('cuda:0', bench.synthetic_inputs())
for item in (('mir', 0), ('mir', 1), ('rot', 0), ('rot', 1), ('rot', 2), ('rot', 3), ('rot', 4)):
    job.step(item)
torch.cuda.synchronize()
torch.cuda.profiler.start()
torch.cuda.nvtx.range_push('timed')
for item in (('mir', 2), ('rot', 0), ('rot', 1), ('rot', 2), ('rot', 3)):
    job.step(item)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
torch.cuda.profiler.stop()
