"""Weight modulation kernels on the generator's layer shapes: forward (demodulation coefficient + modulated weights in the conv
engine's layout) and backward (dW, ds), device time of the kernels from torch.profiler.  Not the headline bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from spi_b200.ops.modulate import modulate_weights

CASES = [  # (N, O, I, k, demod, layout, flip, count per generator pass)
    (1, 512, 512, 3, True, 'ohwi', False, 5),
    (1, 512, 512, 3, True, 'ihwo', True, 4),
    (1, 256, 512, 3, True, 'ihwo', True, 1),
    (1, 256, 256, 3, True, 'ohwi', False, 2),
    (1, 128, 256, 3, True, 'ihwo', True, 2),
    (1, 128, 128, 3, True, 'ohwi', False, 2),
    (1, 96, 128, 1, False, 'ohwi', False, 1),
    (4, 512, 512, 3, True, 'ohwi', False, 0),
]


def kernel_us(fn, reps=5):
    """Device time per call: sum of the CUDA kernel / memset durations torch.profiler records, averaged over `reps` calls."""
    from torch.profiler import ProfilerActivity, profile
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    return sum(e.device_time for e in prof.events() if e.device_type.name == 'CUDA') / reps


def main():
    tot_f = tot_b = 0.0
    for n, o, i, k, demod, layout, flip, cnt in CASES:
        W = torch.randn(o, i, k, k, device='cuda', requires_grad=True)
        s = torch.randn(n, i, device='cuda', requires_grad=True)
        out = modulate_weights(W, s, demod, layout=layout, flip=flip)
        g = torch.ones_like(out)

        def fwd():
            with torch.no_grad():
                modulate_weights(W, s, demod, layout=layout, flip=flip)

        def bwd():
            torch.autograd.grad(out, (W, s), g, retain_graph=True)

        f, b = kernel_us(fwd), kernel_us(bwd)
        mb = o * i * k * k * 4 / 1e6
        print(f'N={n} {o:3d}x{i:3d}x{k}x{k} {layout}{"+flip" if flip else "     "} demod={int(demod)}  W {mb:5.2f} MB   fwd {f:6.1f} us ({(1 + n) * mb / f * 1e3:5.0f} GB/s)   '
              f'bwd {b:6.1f} us ({(2 + n) * mb / b * 1e3:5.0f} GB/s, incl. the ds memset)', flush=True)
        tot_f += cnt * f
        tot_b += cnt * b
    print(f'weighted by layers per generator pass: fwd {tot_f:7.1f} us   bwd {tot_b:7.1f} us (device time, L2-warm)')


if __name__ == '__main__':
    main()
