"""Where do the small ATen launches of an iteration come from?  One eager iteration of each kind under torch.profiler with Python
stacks; prints count / shapes / innermost spi_b200 frame for the element-wise, copy and fill ops.  Not a benchmark."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench

OPS = ('aten::mul', 'aten::mul_', 'aten::add', 'aten::add_', 'aten::copy_', 'aten::fill_', 'aten::zero_', 'aten::sub', 'aten::div', 'aten::neg', 'aten::sum',
       'aten::cat', 'aten::clone', 'aten::contiguous', 'aten::addmm', 'aten::mm', 'aten::zeros', 'aten::sqrt', 'aten::mean', 'aten::where', 'aten::index',
       'aten::rsub', 'aten::pow', 'aten::square', 'aten::sub_', 'aten::div_', 'aten::clamp', 'aten::abs', 'aten::exp', 'aten::stack', 'aten::flip', 'aten::bmm',
       'aten::matmul', 'aten::randn', 'aten::normal_', 'aten::_to_copy', 'aten::lerp', 'aten::sigmoid', 'aten::softplus')


def main():
    from spi_b200.configs import global_config
    global_config.use_cuda_graphs = False
    job = bench.OursJob('cuda:0', bench.synthetic_inputs())
    for item in [('mir', i) for i in range(2)] + [('rot', i) for i in range(4)]:
        job.step(item)
    torch.cuda.synchronize()
    out = open(os.path.join(ROOT, 'gpurun_out', 'op_sites.txt'), 'w')
    for name, kind in (('mir (stage 1)', ('mir', 2)), ('rot light (i%4 != 0)', ('rot', 5)), ('rot heavy (i%4 == 0)', ('rot', 8))):
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
            job.step(kind)
            torch.cuda.synchronize()
        groups = collections.Counter()
        dev = collections.Counter()
        for e in prof.events():
            if e.name not in OPS or e.device_time_total <= 0:
                continue
            # only top-level dispatches: skip ops nested inside another listed op (aten::zeros -> zero_ -> fill_)
            par = e.cpu_parent
            nested = False
            while par is not None:
                if par.name in OPS:
                    nested = True
                    break
                par = par.cpu_parent
            if nested:
                continue
            frames = [s for s in (e.stack or []) if 'spi_b200' in s or 'bench.py' in s]
            site = frames[0].split('/root/repo/')[-1] if frames else '(autograd)'
            bw = ''
            par = e.cpu_parent
            while par is not None:
                if 'Backward' in par.name or 'AccumulateGrad' in par.name:
                    bw = par.name.split(': ')[-1]
                    break
                par = par.cpu_parent
            shapes = str([s for s in (e.input_shapes or []) if s])[:70]
            key = (e.name, site if not bw else bw, shapes)
            groups[key] += 1
            dev[key] += e.device_time_total
        print('=' * 20, name, file=out)
        tot = sum(groups.values())
        print(f'{tot} listed ATen dispatches with device work', file=out)
        for key, c in sorted(groups.items(), key=lambda kv: -kv[1])[:90]:
            print(f'{c:4d}  {dev[key]:8.1f} us  {key[0]:14s} {key[1][:80]:80s} {key[2]}', file=out)
    out.close()
    print(open(os.path.join(ROOT, 'gpurun_out', 'op_sites.txt')).read()[:3000])


if __name__ == '__main__':
    main()
