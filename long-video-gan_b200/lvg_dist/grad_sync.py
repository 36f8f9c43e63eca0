"""Data-parallel gradient exchange for the train_lres.py / train_sres.py step.

The reference synchronises gradients by hand after every G / D / R1 update
(utils.py:104-125): ``torch.cat`` of every ``.grad`` into a fresh flat buffer, one
``all_reduce(SUM)`` per 2**23-element shard, ``/ world_size * gain``, ``nan_to_num(nan=0,
posinf=1e5, neginf=-1e5)``, ``split`` and copy back -- four extra passes over the buffer
around the collective.

Here the flat fp32 buffer is persistent and the parameters' ``.grad`` tensors are views into
it, so backward writes gradients in place; the exchange is ONE NCCL all-reduce over
NVLink 5 / NVSwitch (in-switch NVLS reduction when NCCL enables it) followed by ONE in-place
kernel (``lvg_grad_postprocess``: scale + NaN/Inf clamp). Same results as the reference's
``sync_grads`` (up to the summation order inside NCCL).

``overlap=True`` additionally hides the exchange behind the backward pass: the flat buffer is cut into
buckets in reverse parameter order (the order in which autograd finishes gradients), and a
post-accumulate-grad hook starts the asynchronous all-reduce of a bucket as soon as its last gradient has
been written -- NCCL runs it on its own stream while the remaining backward kernels execute; ``sync()``
then only waits for the tail and runs the post kernel. The reference exchanges after the whole backward.
"""
import ctypes

import torch
import torch.distributed as dist

_GRAD_LIMIT = 1e5     # utils.py:121


def flat_offsets(sizes, align=64):
    """Start offsets (in elements) of tensors of the given sizes in a flat buffer, each on an `align`-element boundary;
    the last entry is the buffer length."""
    offsets, ofs = [], 0
    for n in sizes:
        offsets.append(ofs)
        ofs += (n + align - 1) // align * align
    offsets.append(ofs)
    return offsets


class FlatGradSync:
    """Owns one flat fp32 gradient buffer for a module and exchanges it across ranks.

    >>> sync = FlatGradSync(G)            # once, after building the module
    >>> loss.backward()                   # grads land in sync.flat through the .grad views
    >>> sync.sync(gain=1.0)               # all-reduce mean, sanitise -- replaces utils.sync_grads(G)

    The buffer spans ALL parameters of the module, whatever their ``requires_grad`` flag at construction: the
    reference's loop calls ``sync_grads`` right after ``network.requires_grad_(False)`` (video_gan_lres.py:126-129,
    171-174) and, like ``utils.sync_grads`` (utils.py:116-124), what takes part in an exchange is decided at sync time:
    the parameters whose ``.grad`` is not None. A parameter without a gradient keeps ``.grad is None`` (so Adam skips
    it, as with the reference); its slice of the buffer is exchanged as zeros, which keeps the collective's size
    identical on every rank.

    ``overlap=True``: see the module docstring. A bucket may only leave once EVERY backward pass of the current
    update has written into it -- the reference runs several per sync (``fake_loss.backward()`` then
    ``real_loss.backward()`` in ``update_D``, times ``*_grad_accum``; video_gan_lres.py:135-176). State the number
    with ``backwards_per_sync`` (hooks then count that many gradient arrivals per parameter), or keep the default of
    1 and wrap the LAST backward of an update in ``with sync.final_backward():`` -- outside that context the hooks of
    a ``backwards_per_sync=None`` instance do nothing. A gradient that arrives for a bucket that has already left
    raises (it would silently not be exchanged otherwise). Whatever has not left by ``sync()`` is exchanged there.
    """

    def __init__(self, module, group=None, overlap=False, buckets=4, backwards_per_sync=1):
        self.params = list(module.parameters())
        self.group = group
        self.overlap = bool(overlap)
        self.backwards_per_sync = backwards_per_sync
        self._live = backwards_per_sync is not None
        self._pending, self._works, self._bucket_of, self._slices, self._hooks = [], [], {}, [], []
        if not self.params:
            self.flat = torch.zeros(0)
            return
        device = self.params[0].device
        # every tensor starts on a 256-byte boundary of the flat buffer (vector accesses, and library kernels pick the same
        # code path as for separately allocated tensors); the padding elements stay zero
        self.offsets = flat_offsets([p.numel() for p in self.params])
        self.flat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=device)
        self._views = []
        for p, ofs in zip(self.params, self.offsets):
            if p.dtype != torch.float32:
                raise RuntimeError('FlatGradSync expects fp32 master parameters (as the reference trains)')
            view = self.flat[ofs:ofs + p.numel()].view_as(p)
            if p.requires_grad or p.grad is not None:
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
            self._views.append(view)
        if self.overlap:
            self._make_buckets(max(1, int(buckets)))

    # ---- overlap with the backward pass -------------------------------------------------------------------
    def _make_buckets(self, n_buckets):
        """Contiguous slices of the flat buffer of ~equal size; bucket 0 holds the LAST parameters (ready first)."""
        total = self.flat.numel()
        target = (total + n_buckets - 1) // n_buckets
        start, count = 0, 0
        groups, cur = [], []
        for i, p in enumerate(self.params):
            cur.append(i)
            ofs = self.offsets[i + 1]
            count = ofs - start
            if count >= target or i == len(self.params) - 1:
                groups.append((cur, start, ofs))
                cur, start, count = [], ofs, 0
        groups.reverse()
        self._slices = [(a, b) for _, a, b in groups]
        self._members = [len(idx) for idx, _, _ in groups]
        for b, (idx, _, _) in enumerate(groups):
            for i in idx:
                self._bucket_of[i] = b
        self._arm()
        for i, p in enumerate(self.params):
            was = p.requires_grad                        # hooks can only be registered while the flag is set; they stay
            p.requires_grad_(True)
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
            p.requires_grad_(was)

    def _arm(self):
        k = self.backwards_per_sync or 1
        self._pending = [m * k for m in self._members]
        self._works = [None] * len(self._members)

    def final_backward(self):
        """Context for the last backward pass of an update (``backwards_per_sync=None`` instances): buckets leave as
        soon as that pass has written their last gradient."""
        owner = self

        class _Ctx:
            def __enter__(self):
                owner._pending = list(owner._members)
                owner._live = True

            def __exit__(self, *exc):
                owner._live = False
                return False
        return _Ctx()

    def _make_hook(self, i):
        def hook(param):
            b = self._bucket_of[i]
            view = self._views[i]
            if param.grad is not None and param.grad.data_ptr() != view.data_ptr():
                view.copy_(param.grad)                   # a replaced .grad: fold it back before the bucket leaves
                param.grad = view
            if self._works[b] is not None:
                raise RuntimeError('FlatGradSync(overlap=True): a gradient arrived for a bucket whose all-reduce has already '
                                   'started -- more backward passes per sync() than backwards_per_sync; pass the right count '
                                   'or use final_backward()')
            if not self._live:
                return
            self._pending[b] -= 1
            if self._pending[b] == 0 and self._world() > 1:
                a, e = self._slices[b]
                self._works[b] = dist.all_reduce(self.flat[a:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return hook

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def packed(self):
        """The gradients without the alignment padding, concatenated in parameter order (a copy) -- what
        ``torch.cat([p.grad.flatten() ...])`` of the reference's ``sync_grads`` holds."""
        if not self.params:
            return self.flat.clone()
        return torch.cat([v.reshape(-1) for v in self._views])

    def zero_grad(self):
        """Zero in place (``set_to_none`` would detach the views from the flat buffer)."""
        self.flat.zero_()

    def _reattach(self):
        # optimizers / user code may have replaced p.grad (``zero_grad(set_to_none=True)`` then backward, as the reference
        # loop does): fold such gradients back into the flat buffer with ONE multi-tensor copy; parameters without a
        # gradient stay without one (utils.py:117) and contribute zeros
        src, dst = [], []
        for p, view in zip(self.params, self._views):
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                src.append(p.grad.reshape(view.shape).to(torch.float32))
                dst.append(view)
                p.grad = view
        if src:
            torch._foreach_copy_(dst, src)

    def sync(self, gain=None, postprocess=True):
        """Average the gradients over the process group, scale by `gain` (None = 1, utils.py:120), sanitise NaN/Inf.
        ``postprocess=False`` leaves the summed gradients in the buffer and the outstanding factor in ``pending_scale``:
        ``FlatAdam.step(grad_scale=sync.pending_scale)`` then scales and sanitises inside the optimiser kernel."""
        self.pending_scale = 1.0
        if self.flat.numel() == 0:
            return
        world = self._world()
        if self.overlap:
            # buckets whose hooks all fired are already in flight; anything else (parameters that received no gradient
            # in this update, or fewer backward passes than announced) is exchanged now
            pending = [b for b in range(len(self._slices)) if self._works[b] is None]
            if pending:
                self._reattach()
            for b, (a, e) in enumerate(self._slices):
                if self._works[b] is not None:
                    self._works[b].wait()
                elif world > 1:
                    dist.all_reduce(self.flat[a:e], op=dist.ReduceOp.SUM, group=self.group)
            self._arm()
        else:
            self._reattach()
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if isinstance(gain, torch.Tensor):
            gain = gain.item()
        scale = (1.0 if gain is None else float(gain)) / world
        if postprocess:
            postprocess_(self.flat, scale=scale, limit=_GRAD_LIMIT)
        else:
            self.pending_scale = scale


def postprocess_(flat, scale, limit=_GRAD_LIMIT):
    """In place: ``flat *= scale`` then NaN -> 0, +-inf -> +-limit (``torch.nan_to_num`` semantics)."""
    if flat.is_cuda:
        from torch_utils import custom_ops
        lib = custom_ops.load_library()
        with torch.cuda.device(flat.device):
            rc = lib.lvg_grad_postprocess(ctypes.c_void_p(flat.data_ptr()), flat.numel(), float(scale), float(limit),
                                          ctypes.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream))
        if rc != 0:
            raise RuntimeError('grad_postprocess: ' + lib.lvg_last_error().decode())
    else:
        # CPU tensors only occur in the gloo tests of the host logic
        flat.mul_(scale)
        torch.nan_to_num(flat, nan=0.0, posinf=limit, neginf=-limit, out=flat)
    return flat


def sync_grads(module, gain=None, _cache={}):
    """Drop-in for ``utils.sync_grads(network, gain=None)`` (utils.py:116): keeps one FlatGradSync per module. Works
    with the reference loop as it is: ``requires_grad_(False)`` before the call, ``zero_grad(set_to_none=True)`` after
    the optimiser step (the next backward's fresh ``.grad`` tensors are folded into the buffer by one multi-tensor copy)."""
    key = id(module)
    if key not in _cache:
        _cache[key] = FlatGradSync(module)
    _cache[key].sync(gain)
    return _cache[key]
