"""Flat-buffer optimiser step for the train_lres.py / train_sres.py update.

The reference closes every update with three sweeps over parameter-sized memory, each a few hundred small launches:
``utils.sync_grads`` (scale + ``nan_to_num``, utils.py:116-124), ``torch.optim.Adam.step()`` over ~300 tensors
(model/video_gan_lres.py:84-85,128,174) and, for the generator, ``tensor_ema.lerp_(tensor, 1 - ema_beta)`` over every
parameter and buffer (video_gan_lres.py:208-214). Here parameters, gradients (shared with ``FlatGradSync``) and both
moments live in flat fp32 buffers whose tensors are views, and the whole tail is ONE kernel (``lvg_adam_step``,
csrc/optim.cu): sanitise -> moments -> parameter update -> EMA of the parameters.

``FlatAdam`` keeps the parts of the ``torch.optim.Adam`` surface the reference loop touches: ``param_groups[0]['lr']``
(``update_lrates``), ``step()``, ``zero_grad(set_to_none=True)``, ``state_dict()`` / ``load_state_dict()`` (``ckpt``).
Arithmetic is ``torch.optim.Adam``'s (no weight decay / amsgrad / maximize, the reference's configuration), per-parameter
step counts included: parameters whose ``.grad`` is None in a step are skipped, exactly like torch does; when every
parameter has a gradient (the training loops) the step is a single launch, otherwise one launch per run of consecutive
parameters that share a step count.
"""
import ctypes
import math

import torch


def _lib():
    from torch_utils import custom_ops
    return custom_ops.load_library(), custom_ops


class FlatAdam:
    """Adam over one flat fp32 parameter buffer.

    >>> sync = FlatGradSync(G, overlap=True)          # flat gradients (optional; FlatAdam creates them otherwise)
    >>> opt = FlatAdam(G.parameters(), lr=3e-3, betas=(0, 0.99), grad_sync=sync)
    >>> ema = FlatEMA(G, G_ema, opt)                   # optional: G_ema's parameters updated inside opt.step()
    >>> loss.backward(); sync.sync(gain, postprocess=False)
    >>> opt.step(grad_scale=sync.pending_scale)        # sanitise + Adam + EMA in one kernel

    Parameters become views into ``self.flat_params`` (their values are preserved); gradients are views into
    ``self.flat_grads`` (the ``FlatGradSync`` buffer when one is given -- it must cover the same parameters in the same
    order).
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_sync=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError('FlatAdam: no parameters')
        for p in self.params:
            if p.dtype != torch.float32:
                raise RuntimeError('FlatAdam expects fp32 master parameters (as the reference trains)')
        device = self.params[0].device
        self.param_groups = [dict(params=self.params, lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps))]
        from lvg_dist.grad_sync import flat_offsets
        self._offsets = flat_offsets([p.numel() for p in self.params])      # 256-byte aligned starts, as FlatGradSync lays out
        total = self._offsets[-1]
        self.flat_params = torch.zeros(total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, a in zip(self.params, self._offsets[:-1]):
                view = self.flat_params[a:a + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
        if grad_sync is not None:
            if [id(p) for p in grad_sync.params] != [id(p) for p in self.params]:
                raise ValueError('FlatAdam: grad_sync must cover the same parameters in the same order')
            self.flat_grads = grad_sync.flat
            self._grad_views = grad_sync._views
        else:
            self.flat_grads = torch.zeros(total, dtype=torch.float32, device=device)
            self._grad_views = [self.flat_grads[a:a + p.numel()].view_as(p) for p, a in zip(self.params, self._offsets[:-1])]
            for p, v in zip(self.params, self._grad_views):
                if p.grad is not None:
                    v.copy_(p.grad)
                    p.grad = v
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=device)
        self.steps = [0] * len(self.params)          # per-parameter step counts, as torch.optim.Adam keeps them
        self._ema = None

    # ---- torch.optim surface ------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        """``set_to_none=True`` (the reference's call) detaches ``.grad``; the next backward's fresh tensors are folded back
        into the flat buffer by ``step()`` / ``FlatGradSync``. ``set_to_none=False`` zeroes the flat buffer in place."""
        if set_to_none:
            for p in self.params:
                p.grad = None
        else:
            self.flat_grads.zero_()
            for p, v in zip(self.params, self._grad_views):
                p.grad = v

    def state_dict(self):
        return dict(param_groups=[{k: v for k, v in self.param_groups[0].items() if k != 'params'}], steps=list(self.steps),
                    exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone())

    def load_state_dict(self, state):
        self.param_groups[0].update(state['param_groups'][0])
        self.steps = list(state['steps'])
        self.exp_avg.copy_(state['exp_avg'])
        self.exp_avg_sq.copy_(state['exp_avg_sq'])

    # ---- the step ----------------------------------------------------------------------------------------------
    def _gather(self):
        """Fold gradients that live outside the flat buffer back in; returns the per-parameter 'has a gradient' flags."""
        src, dst, active = [], [], []
        for p, v in zip(self.params, self._grad_views):
            active.append(p.grad is not None)
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad.reshape(v.shape).to(torch.float32))
                dst.append(v)
                p.grad = v
        if src:
            torch._foreach_copy_(dst, src)
        return active

    @torch.no_grad()
    def step(self, grad_scale=None, grad_limit=1e5, ema_beta=None):
        """One Adam step. ``grad_scale`` not None: the gradients are first multiplied by it and sanitised like
        ``utils.sync_grads`` does (NaN -> 0, +-inf -> +-grad_limit) -- pass ``FlatGradSync.pending_scale`` after
        ``sync(..., postprocess=False)``. ``ema_beta`` not None (and a ``FlatEMA`` attached): the EMA of the
        parameters is updated in the same pass."""
        g = self.param_groups[0]
        active = self._gather()
        for i, a in enumerate(active):
            if a:
                self.steps[i] += 1
        ema = self._ema if ema_beta is not None else None
        # runs of consecutive active parameters with equal step counts (ONE run in the training loops)
        runs, i, n = [], 0, len(self.params)
        while i < n:
            if not active[i]:
                i += 1
                continue
            j = i
            while j + 1 < n and active[j + 1] and self.steps[j + 1] == self.steps[i]:
                j += 1
            runs.append((self._offsets[i], self._offsets[j + 1], self.steps[i]))
            i = j + 1
        for a, b, step in runs:
            self._launch(a, b, step, g, grad_scale, grad_limit, ema, ema_beta)
        if ema is not None:
            # parameters skipped by Adam still take part in the average (the reference's loop runs over all of them)
            for i, act in enumerate(active):
                if not act:
                    ema.lerp_range(self._offsets[i], self._offsets[i + 1], ema_beta)
            ema.update_buffers(ema_beta)

    def _launch(self, a, b, step, g, grad_scale, grad_limit, ema, ema_beta):
        p, gr, m, v = self.flat_params[a:b], self.flat_grads[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b]
        pe = ema.flat_params[a:b] if ema is not None else None
        beta1, beta2 = g['betas']
        if p.is_cuda:
            lib, ops = _lib()
            with torch.cuda.device(p.device):
                rc = lib.lvg_adam_step(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), None if pe is None else pe.data_ptr(),
                                       b - a, g['lr'], beta1, beta2, g['eps'], step,
                                       1.0 if grad_scale is None else float(grad_scale),
                                       0.0 if grad_scale is None else float(grad_limit), 1,
                                       0.0 if ema_beta is None else float(ema_beta), ctypes.c_void_p(ops._stream(p)))
            if rc != 0:
                raise RuntimeError('adam_step: ' + lib.lvg_last_error().decode())
            return
        # CPU tensors (host-logic tests): the same arithmetic with torch ops
        if grad_scale is not None:
            gr.mul_(grad_scale)
            torch.nan_to_num(gr, nan=0.0, posinf=grad_limit, neginf=-grad_limit, out=gr)
        m.lerp_(gr, 1 - beta1)
        v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(g['eps'])
        p.addcdiv_(m, denom, value=-g['lr'] / bc1)
        if pe is not None:
            pe.lerp_(p, 1 - ema_beta)


class FlatEMA:
    """Exponential moving average of a network's parameters and buffers (``update_G_ema``, video_gan_lres.py:208-214).

    The averaged parameters live in one flat buffer aligned with ``FlatAdam.flat_params`` and are updated inside
    ``FlatAdam.step(ema_beta=...)``; floating-point buffers are averaged by one ``lvg_lerp`` over a second pair of
    flat buffers, anything else (integer buffers) is copied -- ``lerp_`` with weight 1 - beta on integers is what the
    reference would do as well only for floating types.
    """

    def __init__(self, net, net_ema, optimizer):
        self.opt = optimizer
        params, params_ema = list(net.parameters()), list(net_ema.parameters())
        if [id(p) for p in params] != [id(p) for p in optimizer.params] or len(params) != len(params_ema):
            raise ValueError('FlatEMA: the optimiser must own exactly the parameters of `net`, in order')
        self.flat_params = torch.zeros_like(optimizer.flat_params)
        with torch.no_grad():
            for q, a in zip(params_ema, optimizer._offsets[:-1]):
                view = self.flat_params[a:a + q.numel()].view_as(q)
                view.copy_(q)
                q.data = view
        bufs = [(b, be) for b, be in zip(net.buffers(), net_ema.buffers())]
        self._float = [(b, be) for b, be in bufs if b.is_floating_point() and b.dtype == torch.float32]
        self._other = [(b, be) for b, be in bufs if not (b.is_floating_point() and b.dtype == torch.float32)]
        n = sum(b.numel() for b, _ in self._float)
        device = optimizer.flat_params.device
        self.flat_buf = torch.empty(n, dtype=torch.float32, device=device)          # the live network's buffers, gathered per update
        self.flat_buf_ema = torch.empty(n, dtype=torch.float32, device=device)
        ofs = 0
        with torch.no_grad():
            for b, be in self._float:
                view = self.flat_buf_ema[ofs:ofs + be.numel()].view_as(be)
                view.copy_(be)
                be.data = view
                ofs += be.numel()
        optimizer._ema = self

    @staticmethod
    def _lerp(dst, src, weight):
        if dst.numel() == 0:
            return
        if dst.is_cuda:
            lib, ops = _lib()
            with torch.cuda.device(dst.device):
                rc = lib.lvg_lerp(dst.data_ptr(), src.data_ptr(), dst.numel(), float(weight), ctypes.c_void_p(ops._stream(dst)))
            if rc != 0:
                raise RuntimeError('lerp: ' + lib.lvg_last_error().decode())
        else:
            dst.lerp_(src, float(weight))

    @torch.no_grad()
    def lerp_range(self, a, b, beta):
        """EMA of the parameters in flat range [a, b) outside the fused optimiser kernel."""
        self._lerp(self.flat_params[a:b], self.opt.flat_params[a:b], 1.0 - float(beta))

    @torch.no_grad()
    def update_buffers(self, beta):
        if self._float:
            torch._foreach_copy_(list(self.flat_buf.split([b.numel() for b, _ in self._float])), [b.reshape(-1) for b, _ in self._float])
            self._lerp(self.flat_buf_ema, self.flat_buf, 1.0 - float(beta))
        for b, be in self._other:
            be.copy_(b)


def ema_beta_at(step, ema_beta, warmup_steps):
    """The reference's EMA schedule (video_gan_lres.py:209-210)."""
    reciprocal_halflife = math.log(ema_beta, 0.5) * (warmup_steps + 1) / (step + 1)
    return min(0.5 ** reciprocal_halflife, ema_beta)


class GraphedCallable:
    """CUDA-graph capture of a fixed-shape piece of the step (a network forward+backward on static buffers, an update
    tail): ``fn(*static_inputs)`` is warmed up on a side stream, captured once and then replayed; new inputs are copied
    into the static buffers. Removes the per-launch host cost of the several hundred small kernels of a G/D pass -- what
    the reference pays in eager mode. The callable must not synchronise (no ``.item()``).
    """

    def __init__(self, fn, *static_inputs, warmup=3):
        if not torch.cuda.is_available():
            raise RuntimeError('GraphedCallable needs a CUDA device')
        self.static_inputs = static_inputs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_outputs = fn(*static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if isinstance(dst, torch.Tensor) and src is not dst:
                dst.copy_(src)
        self.graph.replay()
        return self.static_outputs
