"""Achieved HBM GB/s of the streaming kernels of libspi_b200 on the hot shapes of configs[1] (CUDA events, L2 flushed
between iterations).  Bytes are the ALGORITHMIC ones (every operand read once, every result written once).  Not a
benchmark of the product path; run under gpurun (optionally under ncu with -k regex:...)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from spi_b200 import _lib
from spi_b200.torch_utils.ops import bias_act, upfirdn2d

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
QUICK = '--quick' in sys.argv


def time_ms(fn, iters=10):
    if QUICK:
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(f'{name:58s} {nbytes / 1e6:8.1f} MB  {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f}% of measured HBM peak', flush=True)


def cl(*shape):
    return torch.randn(*shape, device='cuda').contiguous(memory_format=torch.channels_last)


def main():
    print(torch.cuda.get_device_name(0), 'peak', PEAK)
    for n, c, h in ((1, 128, 512), (1, 256, 256), (4, 128, 512), (1, 512, 64), (1, 64, 256)):
        x = cl(n, c, h, h)
        b = torch.randn(c, device='cuda')
        nz = torch.randn(h, h, device='cuda')
        st = torch.tensor(0.3, device='cuda')
        nb = x.numel() * 4
        report(f'bias_act_noise fwd lrelu [{n},{c},{h},{h}]', 2 * nb, time_ms(lambda: bias_act.bias_act_noise(x, b, nz, st, act='lrelu', clamp=256)))
        report(f'bias_act fwd relu [{n},{c},{h},{h}]', 2 * nb, time_ms(lambda: bias_act.bias_act(x, b, act='relu')))
        y = bias_act.bias_act_noise(x, b, nz, st, act='lrelu', clamp=256)
        dy = cl(n, c, h, h)
        cfg = (1, bias_act.activation_funcs['lrelu'], 0.2, 2 ** 0.5, 256.0)
        report(f'bias_act grad=1 lrelu [{n},{c},{h},{h}]', 3 * nb, time_ms(lambda: bias_act._BiasActGrad.apply(dy, None, b, y, cfg)))
        report(f'epilogue_grad_reduce [{n},{c},{h},{h}]', nb, time_ms(lambda: bias_act._fused_reductions(dy, True, noise=nz, want_dpix=True, want_ds=True)))
        del x, y, dy
    f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
    for n, c, h in ((1, 128, 513), (1, 256, 257), (4, 128, 513), (4, 256, 257), (1, 128, 257)):
        x = cl(n, c, h, h)
        out = upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4)
        report(f'upfirdn2d blur4 [{n},{c},{h},{h}] -> {tuple(out.shape[2:])}', (x.numel() + out.numel()) * 4,
               time_ms(lambda: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4)))
        b = torch.randn(c, device='cuda')
        nz = torch.randn(h - 1, h - 1, device='cuda')
        st = torch.tensor(0.3, device='cuda')
        with torch.no_grad():
            t2 = time_ms(lambda: bias_act.bias_act(upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4), b, act='lrelu', clamp=256))
            report(f'  blur4 then bias_act (two passes) [{n},{c},{h},{h}]', (x.numel() + out.numel()) * 4, t2)
            report(f'  blur4 + bias + lrelu fused [{n},{c},{h},{h}]', (x.numel() + out.numel()) * 4,
                   time_ms(lambda: bias_act.blur_bias_act_noise(x, f, b, padding=[1, 1, 1, 1], fir_gain=4, act='lrelu', clamp=256)))
            report(f'  blur4 + noise + bias + lrelu fused [{n},{c},{h},{h}]', (x.numel() + out.numel()) * 4,
                   time_ms(lambda: bias_act.blur_bias_act_noise(x, f, b, nz, st, padding=[1, 1, 1, 1], fir_gain=4, act='lrelu', clamp=256)))
        del x, out
    for n, c, h in ((1, 96, 128), (1, 3, 256), (4, 3, 256)):
        x = cl(n, c, h, h)
        out = upfirdn2d.upsample2d(x, f)
        report(f'upsample2d [{n},{c},{h},{h}] -> {tuple(out.shape[2:])}', (x.numel() + out.numel()) * 4, time_ms(lambda: upfirdn2d.upsample2d(x, f)))


if __name__ == '__main__':
    main()
