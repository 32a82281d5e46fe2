import os, sys, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from spi_b200.configs import hyperparameters as hp

which = sys.argv[1]
job = bench.OursJob('cuda:0', bench.synthetic_inputs())
hp.pt_rot_lambda = 0.1 if 'rot' in which else 0
hp.pt_mirror_rot_lambda = 0.05 if 'mir' in which else 0
hp.pt_depth_lambda = 1.0 if 'dep' in which else 0
job.step('mir'); job.step('mir')
print('mir ok', flush=True)
for i in range(6):
    job.i_rot = 0 if i % 2 == 0 else 1
    job.step('rot')
    torch.cuda.synchronize()
    print(which, 'rot step', i, 'ok', flush=True)
