"""Kernel-time breakdown of the bench workload with torch.profiler (device time per kernel name).  Not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench


def main():
    from spi_b200.configs import global_config
    global_config.use_cuda_graphs = '--graphs' in sys.argv
    job = bench.OursJob('cuda:0', bench.synthetic_inputs())
    for item in [('mir', i) for i in range(3)] + [('rot', i) for i in range(4)]:
        job.step(item)
    torch.cuda.synchronize()
    tag = os.environ.get('SPI_CONV_ENGINE', 'tc2')
    for name, kinds in (('mir x4', [('mir', i) for i in range(4)]), ('rot cycle (4 it)', [('rot', i) for i in range(4)])):
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for k in kinds:
                job.step(k)
            torch.cuda.synchronize()
        print('=' * 30, name)
        from torch.autograd import DeviceType
        ev = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA and e.device_time_total > 0]
        tot = sum(e.device_time_total for e in ev)
        ev.sort(key=lambda e: -e.device_time_total)
        print(f'total device time {tot / 1e3:.2f} ms over {len(kinds)} iterations; kernels launched: {sum(e.count for e in ev)}')
        for e in ev[:45]:
            print(f'{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / tot:5.1f}%  n={e.count:5d}  {e.key[:110]}')
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'prof_' + tag + '_' + name.split()[0] + '.txt'), 'w') as f:
            f.write(f'total device time {tot / 1e3:.2f} ms over {len(kinds)} iterations; kernels launched: {sum(e.count for e in ev)}\n')
            for e in ev:
                f.write(f'{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / tot:5.1f}%  n={e.count:5d}  {e.key[:160]}\n')
        import time
        t0 = time.perf_counter()
        for k in kinds:
            job.step(k)
        torch.cuda.synchronize()
        print(f'wall (unprofiled) {1e3 * (time.perf_counter() - t0) / len(kinds):.2f} ms / iteration')


if __name__ == '__main__':
    main()
