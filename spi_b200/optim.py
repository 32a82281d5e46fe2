"""Fused Adam over flat arenas (replaces `torch.optim.Adam` at spi/training/coaches/base_coach.py:132-135 and
spi/training/projectors/*_projector.py:55-58; defaults betas=(0.9,0.999), eps=1e-8).

Every parameter is re-seated as a view into one contiguous fp32 arena and its `.grad` as a view into a matching
gradient arena, so `zero_grad()` is one memset and `step()` one `spi_adam_step` launch streaming 28 B/param.
`param_groups[0]['lr']` is honoured at step time (the projectors rewrite it every step, mirror_projector.py:88-91).
Parameters that never receive a gradient keep a zero gradient and therefore never move (m = v = 0 -> update 0), which
matches torch skipping `grad is None` parameters.
"""
import torch

from . import _lib


class FlatAdam:
    CHUNK = 32768          # elements per table row of the multi-tensor step (one CTA each)
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, steal_grads=False):
        self.steal = bool(steal_grads)     # leave .grad to autograd (no gradient arena, no per-parameter accumulation launch)
        self.params = [p for p in params]
        assert self.params, 'FlatAdam: empty parameter list'
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('spi_b200.FlatAdam: parameters must reside on a CUDA device (no CPU path in this build)')
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]       # keep every view 16-byte aligned
        total = sum(sizes)
        self.arena = torch.zeros(total, device=dev)
        self.grads = torch.zeros(0 if self.steal else total, device=dev)
        self._offsets, off_ = [], 0
        for sz_ in sizes:
            self._offsets.append(off_)
            off_ += sz_
        self._max_rows = sum((p.numel() + self.CHUNK - 1) // self.CHUNK for p in self.params)
        self._tables = []            # (pinned host table, device table) pairs; one per captured graph + one reused in eager mode
        self._eager_table = None
        self._keep = []
        self.exp_avg = torch.zeros(total, device=dev)
        self.exp_avg_sq = torch.zeros(total, device=dev)
        off = 0
        with torch.no_grad():
            for p, sz in zip(self.params, sizes):
                view = self.arena[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = None if self.steal else self.grads[off:off + p.numel()].view(p.shape)
                off += sz
        self.param_groups = [dict(params=self.params, lr=lr, betas=betas, eps=eps)]
        self.steps = 0
        self.hyper = None            # device {lr, bc1, bc2}: enable with use_device_hyper() for CUDA-graph replay
        self.skip_if_le = None       # (device scalar tensor, threshold): device-side early exit

    def flat_grads(self):
        """Gradients of all parameters as one flat tensor in arena order (zeros where a parameter received none)."""
        if not self.steal:
            return self.grads.clone()
        out = torch.zeros_like(self.arena)
        for p, off in zip(self.params, self._offsets):
            if p.grad is not None:
                out[off:off + p.numel()] = p.grad.reshape(-1)
        return out

    def zero_grad(self, set_to_none=False):
        if self.steal:
            for p in self.params:
                p.grad = None
            return
        self.grads.zero_()
        for p in self.params:           # autograd may have replaced .grad (e.g. after set_to_none elsewhere): re-seat
            if p.grad is None or p.grad.data_ptr() < self.grads.data_ptr() or p.grad.data_ptr() >= self.grads.data_ptr() + self.grads.numel() * 4:
                self._reseat()
                break

    def _reseat(self):
        if self.steal:
            return
        off = 0
        for p in self.params:
            sz = (p.numel() + 3) // 4 * 4
            view = self.grads[off:off + p.numel()].view(p.shape)
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
            off += sz

    def use_device_hyper(self):
        """Keep {lr, 1-b1^t, 1-b2^t} in a device tensor refreshed by `advance()` so that `step()` can live inside a CUDA graph."""
        self.hyper = torch.zeros(3, device=self.arena.device)

    def advance(self):
        """Host side of a graphed step: bump the step counter and upload the scalars the captured kernel reads."""
        self.steps += 1
        g = self.param_groups[0]
        # fill_ passes the value as a kernel argument fixed at enqueue time.  (An async copy from ONE reused pinned buffer reads
        # the buffer when the copy executes: the host, running hundreds of graph replays ahead, would have overwritten it.)
        self.hyper[0:1].fill_(float(g['lr']))
        self.hyper[1:2].fill_(1.0 - g['betas'][0] ** self.steps)
        self.hyper[2:3].fill_(1.0 - g['betas'][1] ** self.steps)

    @torch.no_grad()
    def step(self, in_graph=False):
        if not in_graph:
            self._reseat()
            if self.hyper is None:
                self.steps += 1
            else:
                self.advance()
        g = self.param_groups[0]
        cond, thr = self.skip_if_le if self.skip_if_le is not None else (None, 0.0)
        if self.steal:
            return self._step_multi(g, cond, thr)
        with _lib.timed('adam', self.arena.numel() * 28):
            _lib.check(_lib.load().spi_adam_step(_lib.ptr(self.arena), _lib.ptr(self.grads), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                                 self.arena.numel(), float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']),
                                                 max(self.steps, 1), _lib.ptr(self.hyper) if self.hyper is not None else None, 0,
                                                 _lib.ptr(cond) if cond is not None else None, float(thr), _lib.stream()))

    def _step_multi(self, g, cond, thr):
        """One launch over a pointer table {param, grad, m, v, numel} of the parameters that received a gradient."""
        rows, self._keep = [], []
        for p, off in zip(self.params, self._offsets):
            gr = p.grad
            if gr is None:
                continue
            if gr.dtype != torch.float32 or not gr.is_contiguous() or gr.data_ptr() % 4:
                gr = gr.contiguous().float()
                self._keep.append(gr)
            # rows are chunks of at most CHUNK elements so that every CTA has the same amount of work
            for c0 in range(0, p.numel(), self.CHUNK):
                base = (off + c0) * 4
                rows.append([self.arena.data_ptr() + base, gr.data_ptr() + c0 * 4, self.exp_avg.data_ptr() + base, self.exp_avg_sq.data_ptr() + base,
                             min(self.CHUNK, p.numel() - c0)])
        if not rows:
            return
        capturing = torch.cuda.is_current_stream_capturing()
        new_pair = lambda: (torch.empty(self._max_rows, 5, dtype=torch.int64).pin_memory(),
                            torch.empty(self._max_rows, 5, dtype=torch.int64, device=self.arena.device))
        if capturing:
            # a captured graph re-uploads ITS table on every replay: it must own the pinned buffer (eager steps in between
            # would otherwise overwrite the pointers the graph was captured with)
            host, dev = new_pair()
            self._tables.append((host, dev))
            host[:len(rows)] = torch.tensor(rows, dtype=torch.int64)
            dev.copy_(host, non_blocking=True)
        else:
            # eager: a ring of pinned slots, each guarded by an event recorded after the kernel that read its device table, so
            # a slot is never rewritten while its upload (or the step reading it) is still in flight
            if self._eager_table is None:
                self._eager_table = {'slots': [], 'next': 0}
            ring = self._eager_table
            if len(ring['slots']) < 4:
                ring['slots'].append([*new_pair(), None])
                slot = ring['slots'][-1]
            else:
                slot = ring['slots'][ring['next'] % 4]
                ring['next'] += 1
                if slot[2] is not None:
                    slot[2].synchronize()
            host, dev = slot[0], slot[1]
            host[:len(rows)] = torch.tensor(rows, dtype=torch.int64)
            dev.copy_(host, non_blocking=True)
        nbytes = sum(r[4] for r in rows) * 28
        with _lib.timed('adam', nbytes):
            _lib.check(_lib.load().spi_adam_step_multi(_lib.ptr(dev), len(rows), 1, float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']),
                                                       max(self.steps, 1), _lib.ptr(self.hyper) if self.hyper is not None else None,
                                                       _lib.ptr(cond) if cond is not None else None, float(thr), _lib.stream()))
        if not capturing:
            slot[2] = torch.cuda.Event()
            slot[2].record()
