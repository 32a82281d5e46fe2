"""Importance-sampled tri-plane volume renderer (drop-in for `training.volumetric_rendering.renderer`,
renderer.py:84-253).

`ImportanceRenderer.forward(planes, decoder, ray_origins, ray_directions, rendering_options)` keeps the reference
signature and return value; the whole chain (stratified sampling, tri-plane gather, OSG decoder, coarse march,
importance sampling, merge, final march) runs in `spi_render_forward` / `spi_render_backward`
(spi_b200/csrc/raymarch.cu).  Randomness: the two uniform draws of renderer.py:190,237 are taken from torch's CUDA
generator with the reference's shapes and order; tests inject them through `ImportanceRenderer.inject_noise`.
"""
import torch

from ... import _lib
from .ray_marcher import MipRayMarcher2

SCRATCH_COLS = (32, 64, 64, 36)
PER_VIEW_GRADIENT_PLANES = False   # shared planes, n views: one gradient plane set per view, summed after the kernel (measured: no gain,
                                   # the views' REDs do not contend noticeably; kept as a switch)
KEEP_ACTIVATIONS = True       # forward keeps hidden layer / outputs / features for backward when decoder gradients are wanted (tcgen05 kernels)
KERNEL_TIMER = None      # bench.py installs an object with .start(tag) / .stop(tag) that records CUDA events on the launch stream


def generate_planes():
    """renderer.py:22-37 (kept for API parity; the kernel hard-wires these three projections)."""
    return torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                         [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
                         [[0, 0, 1], [1, 0, 0], [0, 1, 0]]], dtype=torch.float32)


def _planes_nhwc(planes):
    """[N,3,32,H,W] (any strides) -> channels-last storage [N,H,W,96] as an [N,96,H,W] tensor."""
    n, p, c, h, w = planes.shape
    assert p == 3 and c == 32, 'tri-plane features must be [N, 3, 32, H, W]'
    return planes.reshape(n, p * c, h, w).contiguous(memory_format=torch.channels_last)


def _decoder_tensors(decoder):
    fc1, fc2 = decoder.net[0], decoder.net[2]
    lr_mul = float(fc1.bias_gain)
    assert fc1.weight.shape == (64, 32) and fc2.weight.shape == (33, 64), 'OSGDecoder must be 32 -> 64 -> 33'
    return fc1.weight, fc1.bias, fc2.weight, fc2.bias, lr_mul


def _decoder_grads(sc, n_rows, lr_mul):
    """dW1 = dpre^T f, db1 = sum dpre, dW2 = dout^T hid, db2 = sum dout (gains of FullyConnectedLayer folded back)."""
    f, hid, dpre, dout = sc
    g1 = lr_mul / (32 ** 0.5)
    g2 = lr_mul / (64 ** 0.5)
    # K = millions of samples, outputs 64x32 / 36x64: the convolution weight-gradient kernel with a 1x1 "tap" is exactly this reduction
    # (out[m][c] = sum_rows u[row][m] v[row][c]: TF32 operands through TMA, fp32 accumulation in tensor memory, the row range split over the
    # SMs, reduce-add stores), and its idle epilogue warps add up the columns of u on the way: weight and bias gradients of a layer come
    # out of ONE pass that reads every scratch byte once (spi_rows_outer_sum; before: two batched GEMMs over 512 row chunks, two sums over
    # the chunks and two column-sum passes that re-read d_pre / d_out)
    lib = _lib.load()
    dev = f.device
    dw1, s1 = torch.empty(64, 32, device=dev), torch.empty(64, device=dev)
    dw2, s2 = torch.empty(36, 64, device=dev), torch.empty(36, device=dev)
    if n_rows % 8 == 0 and n_rows >= 8:
        _lib.check(lib.spi_rows_outer_sum(_lib.ptr(dpre), _lib.ptr(f), n_rows, 64, 32, _lib.ptr(dw1), _lib.ptr(s1), _lib.stream()))
        _lib.check(lib.spi_rows_outer_sum(_lib.ptr(dout), _lib.ptr(hid), n_rows, 36, 64, _lib.ptr(dw2), _lib.ptr(s2), _lib.stream()))
    else:           # ragged row counts (stand-alone point queries): library GEMMs + the streaming column sums
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            dw1 = dpre[:n_rows].t() @ f[:n_rows]
            dw2 = dout[:n_rows].t() @ hid[:n_rows]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        _lib.check(lib.spi_column_sums(_lib.ptr(dpre), n_rows, 64, _lib.ptr(s1), _lib.stream()))
        _lib.check(lib.spi_column_sums(_lib.ptr(dout), n_rows, 36, _lib.ptr(s2), _lib.stream()))
    dw1 = dw1 * g1
    dw2 = dw2[:33] * g2
    db1 = s1 * lr_mul
    db2 = s2[:33] * lr_mul
    return dw1, db1, dw2, db2


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, planes, w1, b1, w2, b2, origins, dirs, jitter, u, opts):
        lib = _lib.load()
        np_, _, h, w = planes.shape
        n, r = origins.shape[0], origins.shape[1]
        assert np_ in (1, n), 'planes batch must be 1 (shared by all views) or match the number of cameras'
        plane_bs = 0 if (np_ == 1 and n > 1) else h * w * 96
        dc, df = opts['dc'], opts['df']
        dev = planes.device
        feat = torch.empty(n, r, 32, device=dev)
        depth = torch.empty(n, r, 1, device=dev)
        wsum = torch.empty(n, r, 1, device=dev)
        depths_all = torch.empty(n, r, dc + df, device=dev)
        minmax = torch.empty(2, dtype=torch.int32, device=dev)
        # Keeping activations pays when the decoder is being trained (the backward pass would otherwise write the same f / hid rows
        # for the weight-gradient GEMMs: measured 0.31 + 1.31 -> 0.42 + 1.10 ms/img); with frozen decoder weights the forward
        # pass's extra 0.4 GB/img of stores cost what the backward pass saves (its floor is the L2 atomic traffic of the scatter)
        keep = any(ctx.needs_input_grad[1:5]) and KEEP_ACTIVATIONS and bool(lib.spi_render_keeps_activations(dc, df))
        kept = None
        if keep:
            # the forward pass keeps hidden layer / outputs (/ features) of every sample for the backward pass: ~0.5 GB per image
            rows = n * r * (dc + df)
            need_dec = any(ctx.needs_input_grad[1:5])
            kept = (torch.empty(rows, 64, device=dev), torch.empty(rows, 36, device=dev),
                    torch.empty(rows, 32, device=dev) if need_dec else None, torch.empty(n, r, dc + df, dtype=torch.uint8, device=dev))
        if KERNEL_TIMER is not None:
            KERNEL_TIMER.start('render_fwd', n)
        if keep:
            _lib.check(lib.spi_render_forward_keep(
                _lib.ptr(planes), _lib.ptr(origins), _lib.ptr(dirs), _lib.ptr(jitter), _lib.ptr(u), _lib.ptr(w1), _lib.ptr(b1),
                _lib.ptr(w2), _lib.ptr(b2), opts['lr_mul'], _lib.ptr(feat), _lib.ptr(depth), _lib.ptr(wsum), _lib.ptr(depths_all),
                _lib.ptr(minmax), n, r, plane_bs, h, w, dc, df, opts['ray_start'], opts['ray_end'], opts['box_warp'],
                int(opts['disparity']), _lib.ptr(kept[0]), _lib.ptr(kept[1]), _lib.ptr(kept[2]), _lib.ptr(kept[3]), _lib.stream()))
        else:
            _lib.check(lib.spi_render_forward(
                _lib.ptr(planes), _lib.ptr(origins), _lib.ptr(dirs), _lib.ptr(jitter), _lib.ptr(u), _lib.ptr(w1), _lib.ptr(b1),
                _lib.ptr(w2), _lib.ptr(b2), opts['lr_mul'], _lib.ptr(feat), _lib.ptr(depth), _lib.ptr(wsum), _lib.ptr(depths_all),
                None, _lib.ptr(minmax), n, r, plane_bs, h, w, dc, df, opts['ray_start'], opts['ray_end'], opts['box_warp'],
                int(opts['disparity']), _lib.stream()))
        if KERNEL_TIMER is not None:
            KERNEL_TIMER.stop('render_fwd')
        ctx.save_for_backward(planes, w1, b1, w2, b2, origins, dirs, depths_all, minmax)
        ctx.opts, ctx.plane_bs, ctx.kept = opts, plane_bs, kept
        ctx.mark_non_differentiable(wsum)
        return feat, depth, wsum

    @staticmethod
    def backward(ctx, g_feat, g_depth, _g_wsum):
        planes, w1, b1, w2, b2, origins, dirs, depths_all, minmax = ctx.saved_tensors
        opts = ctx.opts
        lib = _lib.load()
        _, _, h, w = planes.shape
        n, r = origins.shape[0], origins.shape[1]
        plane_bs = ctx.plane_bs
        dc, df = opts['dc'], opts['df']
        need_planes = ctx.needs_input_grad[0]
        need_dec = any(ctx.needs_input_grad[1:5])
        g_feat = g_feat.contiguous()
        g_depth = g_depth.contiguous() if g_depth is not None else None
        # channels-last gradient arena, accumulated with RED (optionally one plane set per view, see PER_VIEW_GRADIENT_PLANES)
        per_view = need_planes and plane_bs == 0 and n > 1 and PER_VIEW_GRADIENT_PLANES
        gplane_bs = h * w * 96 if per_view else plane_bs
        g_planes = None
        if need_planes:
            g_planes = (torch.empty(n, *planes.shape[1:], device=planes.device, memory_format=torch.channels_last).zero_() if per_view
                        else torch.zeros_like(planes))
        gw = None
        if KERNEL_TIMER is not None:
            KERNEL_TIMER.start('render_bwd', n)
        kept = ctx.kept
        if kept is not None:
            rows = n * r * (dc + df)
            sc_dpre = torch.empty(rows, 64, device=planes.device) if need_dec else None
            sc_dout = torch.empty(rows, 36, device=planes.device) if need_dec else None
            _lib.check(lib.spi_render_backward_kept(
                _lib.ptr(planes), _lib.ptr(origins), _lib.ptr(dirs), _lib.ptr(depths_all), _lib.ptr(minmax), _lib.ptr(w1),
                _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), opts['lr_mul'], _lib.ptr(g_feat), _lib.ptr(g_depth),
                _lib.ptr(g_planes), _lib.ptr(sc_dpre), _lib.ptr(sc_dout), n, r, plane_bs, gplane_bs, h, w, dc, df, opts['box_warp'],
                _lib.ptr(kept[0]), _lib.ptr(kept[1]), _lib.ptr(kept[3]), _lib.stream()))
            if KERNEL_TIMER is not None:
                KERNEL_TIMER.stop('render_bwd')
            gw = (None, None, None, None)
            if need_dec:
                if KERNEL_TIMER is not None:
                    KERNEL_TIMER.start('render_dec_grads', n)
                gw = _decoder_grads((kept[2], kept[0], sc_dpre, sc_dout), rows, opts['lr_mul'])
                if KERNEL_TIMER is not None:
                    KERNEL_TIMER.stop('render_dec_grads')
        elif need_dec:
            # all images in one launch: the per-sample rows (784 B/sample) of the whole batch feed two TF32 GEMMs
            rows = n * r * (dc + df)
            sc = [torch.empty(rows, c, device=planes.device) for c in SCRATCH_COLS]
            _lib.check(lib.spi_render_backward(
                _lib.ptr(planes), _lib.ptr(origins), _lib.ptr(dirs), _lib.ptr(depths_all), _lib.ptr(minmax), _lib.ptr(w1),
                _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), opts['lr_mul'], _lib.ptr(g_feat), _lib.ptr(g_depth),
                _lib.ptr(g_planes), _lib.ptr(sc[0]), _lib.ptr(sc[1]), _lib.ptr(sc[2]), _lib.ptr(sc[3]), n, r, plane_bs, gplane_bs, h, w, dc, df,
                opts['box_warp'], _lib.stream()))
            if KERNEL_TIMER is not None:
                KERNEL_TIMER.stop('render_bwd')
                KERNEL_TIMER.start('render_dec_grads', n)
            gw = _decoder_grads(sc, rows, opts['lr_mul'])
            if KERNEL_TIMER is not None:
                KERNEL_TIMER.stop('render_dec_grads')
        else:
            _lib.check(lib.spi_render_backward(
                _lib.ptr(planes), _lib.ptr(origins), _lib.ptr(dirs), _lib.ptr(depths_all), _lib.ptr(minmax), _lib.ptr(w1),
                _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), opts['lr_mul'], _lib.ptr(g_feat), _lib.ptr(g_depth),
                _lib.ptr(g_planes), None, None, None, None, n, r, plane_bs, gplane_bs, h, w, dc, df, opts['box_warp'], _lib.stream()))
            gw = (None, None, None, None)
            if KERNEL_TIMER is not None:
                KERNEL_TIMER.stop('render_bwd')
        if per_view:
            g_planes = g_planes.sum(0, keepdim=True)
        return (g_planes, *gw, None, None, None, None, None)


class _PointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, planes, w1, b1, w2, b2, coords, opts):
        n, _, h, w = planes.shape
        m = coords.shape[1]
        rgb = torch.empty(n, m, 32, device=planes.device)
        sigma = torch.empty(n, m, 1, device=planes.device)
        _lib.check(_lib.load().spi_points_forward(
            _lib.ptr(planes), _lib.ptr(coords), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), opts['lr_mul'],
            _lib.ptr(rgb), _lib.ptr(sigma), n, m, h, w, opts['box_warp'], _lib.stream()))
        ctx.save_for_backward(planes, w1, b1, w2, b2, coords)
        ctx.opts = opts
        return rgb, sigma

    @staticmethod
    def backward(ctx, g_rgb, g_sigma):
        planes, w1, b1, w2, b2, coords = ctx.saved_tensors
        opts = ctx.opts
        n, _, h, w = planes.shape
        m = coords.shape[1]
        need_planes = ctx.needs_input_grad[0]
        need_dec = any(ctx.needs_input_grad[1:5])
        g_planes = torch.zeros_like(planes) if need_planes else None
        sc = [torch.empty(n * m, c, device=planes.device) for c in SCRATCH_COLS] if need_dec else [None] * 4
        _lib.check(_lib.load().spi_points_backward(
            _lib.ptr(planes), _lib.ptr(coords), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), opts['lr_mul'],
            _lib.ptr(g_rgb.contiguous()) if g_rgb is not None else None,
            _lib.ptr(g_sigma.contiguous()) if g_sigma is not None else None, _lib.ptr(g_planes),
            _lib.ptr(sc[0]), _lib.ptr(sc[1]), _lib.ptr(sc[2]), _lib.ptr(sc[3]), n, m, h, w, opts['box_warp'], _lib.stream()))
        gw = _decoder_grads(sc, n * m, opts['lr_mul']) if need_dec else (None,) * 4
        return (g_planes, *gw, None, None)


class ImportanceRenderer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ray_marcher = MipRayMarcher2()
        self.plane_axes = generate_planes()
        self._noise_queue = []

    def inject_noise(self, jitter, u):
        """Queue the (jitter [N,R,Dc,1], u [N*R,Df]) pair the next forward() consumes instead of drawing from the RNG."""
        self._noise_queue.append((jitter, u))

    @staticmethod
    def _check_options(o):
        if isinstance(o['ray_start'], str) or isinstance(o['ray_end'], str):
            raise NotImplementedError("ray_start/ray_end='auto' (per-ray box limits, renderer.py:91-97) is not built: the "
                                      'FFHQ generator of the inversion path uses scalar limits')
        if o.get('density_noise', 0) > 0:
            raise NotImplementedError('density_noise > 0 (renderer.py:146-147) is a GAN-training feature, not built')
        if o.get('white_back', False):
            raise NotImplementedError('white_back=True is not used by the FFHQ generator, not built')
        assert o.get('clamp_mode', 'softplus') == 'softplus', "MipRayMarcher only supports `clamp_mode`=`softplus`!"

    def forward(self, planes, decoder, ray_origins, ray_directions, rendering_options):
        o = rendering_options
        self._check_options(o)
        if not planes.is_cuda:
            raise RuntimeError('spi_b200.ImportanceRenderer: tensors must reside on a CUDA device (no CPU path in this build)')
        n, r, _ = ray_origins.shape
        dc, df = int(o['depth_resolution']), int(o['depth_resolution_importance'])
        if self._noise_queue:
            jitter, u = self._noise_queue.pop(0)
        else:   # same shapes and order as renderer.py:190 (rand_like of [N,R,Dc,1]) and :237 (rand [N*R, Df])
            jitter = torch.rand(n, r, dc, 1, device=planes.device)
            u = torch.rand(n * r, max(df, 1), device=planes.device)
        w1, b1, w2, b2, lr_mul = _decoder_tensors(decoder)
        opts = dict(dc=dc, df=df, lr_mul=lr_mul, ray_start=float(o['ray_start']), ray_end=float(o['ray_end']),
                    box_warp=float(o['box_warp']), disparity=bool(o.get('disparity_space_sampling', False)))
        feat, depth, wsum = _RenderFn.apply(_planes_nhwc(planes), w1, b1, w2, b2, ray_origins.detach().float().contiguous(),
                                            ray_directions.detach().float().contiguous(), jitter.float().contiguous(),
                                            u.float().contiguous(), opts)
        return feat, depth, wsum

    def run_model(self, planes, decoder, sample_coordinates, sample_directions, options):
        """renderer.py:142-149: {'rgb': [N,M,32], 'sigma': [N,M,1]} at arbitrary points."""
        if options.get('density_noise', 0) > 0:
            raise NotImplementedError('density_noise > 0 is not built')
        w1, b1, w2, b2, lr_mul = _decoder_tensors(decoder)
        opts = dict(lr_mul=lr_mul, box_warp=float(options['box_warp']))
        rgb, sigma = _PointsFn.apply(_planes_nhwc(planes), w1, b1, w2, b2, sample_coordinates.detach().float().contiguous(), opts)
        return {'rgb': rgb, 'sigma': sigma}

    # ---- stand-alone stages, same names as the reference methods (forward only; index outputs are int64) ----
    def sample_importance(self, z_vals, weights, N_importance, u=None):
        """renderer.py:194-212; `u` optionally injects the uniform draw of sample_pdf (:237)."""
        n, r, dc, _ = z_vals.shape
        z = z_vals.detach().reshape(n * r, dc).float().contiguous()
        w = weights.detach().reshape(n * r, -1).float().contiguous()
        if u is None:
            u = torch.rand(n * r, N_importance, device=z.device)
        fine = torch.empty(n * r, N_importance, device=z.device)
        inds = torch.empty(n * r, N_importance, dtype=torch.int32, device=z.device)
        _lib.check(_lib.load().spi_sample_importance(_lib.ptr(z), _lib.ptr(w), _lib.ptr(u.contiguous()), n * r, dc, N_importance,
                                                     _lib.ptr(fine), _lib.ptr(inds), None, _lib.stream()))
        return fine.reshape(n, r, N_importance, 1)

    def sample_pdf_from_cdf(self, bins, cdf, u):
        """Inverse-CDF half of sample_pdf (renderer.py:241-253): returns (samples, inds int64)."""
        rays, ncdf = cdf.shape
        df = u.shape[1]
        fine = torch.empty(rays, df, device=cdf.device)
        inds = torch.empty(rays, df, dtype=torch.int32, device=cdf.device)
        _lib.check(_lib.load().spi_inverse_cdf(_lib.ptr(bins.float().contiguous()), _lib.ptr(cdf.float().contiguous()),
                                               _lib.ptr(u.float().contiguous()), rays, ncdf, bins.shape[1], df, _lib.ptr(fine),
                                               _lib.ptr(inds), _lib.stream()))
        return fine, inds.long()

    def unify_samples(self, depths1, colors1, densities1, depths2, colors2, densities2):
        """renderer.py:157-167 (forward only): merge by depth."""
        n, r, dc, _ = depths1.shape
        df = depths2.shape[2]
        perm = torch.empty(n * r, dc + df, dtype=torch.int32, device=depths1.device)
        srt = torch.empty(n * r, dc + df, device=depths1.device)
        _lib.check(_lib.load().spi_unify_samples(_lib.ptr(depths1.detach().float().reshape(n * r, dc).contiguous()),
                                                 _lib.ptr(depths2.detach().float().reshape(n * r, df).contiguous()), n * r, dc, df,
                                                 _lib.ptr(perm), _lib.ptr(srt), _lib.stream()))
        idx = perm.long().reshape(n, r, dc + df, 1)
        all_colors = torch.gather(torch.cat([colors1, colors2], -2), -2, idx.expand(-1, -1, -1, colors1.shape[-1]))
        all_dens = torch.gather(torch.cat([densities1, densities2], -2), -2, idx)
        return srt.reshape(n, r, dc + df, 1), all_colors, all_dens

    def sort_permutation(self, depths1, depths2):
        n, r, dc, _ = depths1.shape
        df = depths2.shape[2]
        perm = torch.empty(n * r, dc + df, dtype=torch.int32, device=depths1.device)
        srt = torch.empty(n * r, dc + df, device=depths1.device)
        _lib.check(_lib.load().spi_unify_samples(_lib.ptr(depths1.detach().float().reshape(n * r, dc).contiguous()),
                                                 _lib.ptr(depths2.detach().float().reshape(n * r, df).contiguous()), n * r, dc, df,
                                                 _lib.ptr(perm), _lib.ptr(srt), _lib.stream()))
        return perm.long().reshape(n, r, dc + df, 1), srt.reshape(n, r, dc + df, 1)
