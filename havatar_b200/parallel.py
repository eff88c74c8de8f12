"""Data-parallel gradient exchange for the two training steps: one process per GPU, frames sharded
across ranks, and ONE exchange step per optimiser step -- all-reduce(SUM)/world on the gradients only.  The reference has no
multi-GPU render path of its own (train_avatarHD.py wraps its modules in torch DistributedDataParallel, :125-134); this is the
B200-side equivalent, written for NVLink 5 / NVSwitch:

  * gradients live in a few large flat fp32 buckets (parameters' .grad are views into them), so one collective moves
    64 MiB instead of hundreds of small tensors -- on NVSwitch the cost is launch latency + bytes / bus bandwidth, not
    per-link hops, so buckets are sized for overlap granularity only;
  * buckets are filled in reverse parameter order (the order backward produces gradients); a post-accumulate hook counts a
    bucket's parameters down and issues its all-reduce asynchronously the moment the last one lands, so the exchange of the
    early buckets overlaps the rest of backward (NCCL runs it on its own stream);
  * collectives are issued in the order the buckets complete, which is the same on every rank (same autograd graph), so ranks can
    never disagree on the order of NCCL calls;
  * `finish()` issues what is left (parameters that received no gradient this step), waits, and divides by the world size
    (ReduceOp.AVG inside NCCL; an explicit scale on backends without it, e.g. gloo in the CPU tests).

Nothing here touches the render data path: rays / frames shard with no collective (havatar_b200/shard.py)."""
import torch
import torch.distributed as dist


def broadcast_parameters(modules, src=0, group=None):
    """Same initial weights (and buffers) on every rank: broadcast rank `src`'s copies, flattened per dtype."""
    if not isinstance(modules, (list, tuple)):
        modules = [modules]
    tensors = []
    for m in modules:
        tensors += [p.data for p in m.parameters()] + [b.data for b in m.buffers()]
    by_type = {}
    for t in tensors:
        by_type.setdefault((t.dtype, t.device), []).append(t)
    for ts in by_type.values():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


class FlatLayout:
    """The parameters of one optimiser laid out in ONE flat fp32 buffer, with a gradient buffer of the same shape: every
    `p.data` and `p.grad` becomes a view at the same (128-byte aligned) offset.  autograd accumulates into the views in place,
    the gradient exchange all-reduces contiguous slices, and the optimiser step is one pass over the buffers (FlatAdam)."""

    def __init__(self, params):
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        if not uniq:
            raise ValueError("FlatLayout needs at least one parameter")
        dev, dt = uniq[0].device, uniq[0].dtype
        if any(p.device != dev or p.dtype != dt for p in uniq):
            raise ValueError("FlatLayout: parameters must share one device and dtype")
        self.offsets, off = [], 0
        for p in uniq:
            self.offsets.append(off)
            off += -(-p.numel() // 32) * 32
        self.numel = off
        self.flat_p = torch.zeros(off, dtype=dt, device=dev)
        self.flat_g = torch.zeros(off, dtype=dt, device=dev)
        with torch.no_grad():
            for p, o in zip(uniq, self.offsets):
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)

    def check(self):
        lo, hi = self.flat_g.data_ptr(), self.flat_g.data_ptr() + self.flat_g.numel() * self.flat_g.element_size()
        for p in self.params:
            if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                raise RuntimeError("a parameter's .grad no longer aliases the flat gradient buffer; use the owner's zero_grad()")


class FlatAdam:
    """torch.optim.Adam semantics (amsgrad=False, no weight decay) as ONE kernel over a FlatLayout (C ABI: hav_adam_flat): reads
    p, g, m, v and writes p, m, v and g = 0 -- 32 B per parameter at HBM speed instead of ~14 multi-tensor launches, and no
    separate zero_grad pass.  Step count and learning rate live on the device (CUDA-graph capturable; set_lr between replays).
    Difference to torch: a parameter that received no gradient in a step is treated as having a zero gradient (torch skips
    parameters whose .grad is None) -- identical whenever gradients are zeroed rather than dropped."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, layout=None):
        self.layout = layout if layout is not None else FlatLayout(params)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.m, self.v = torch.zeros_like(self.layout.flat_p), torch.zeros_like(self.layout.flat_p)
        self.state = torch.tensor([0.0, float(lr)], dtype=torch.float32, device=self.layout.flat_p.device)

    def set_lr(self, lr):
        self.state[1:2].fill_(float(lr))

    def step(self, grad_scale=1.0):
        import ctypes as C

        from . import _lib, styleunet

        L, lay = _lib.lib(), self.layout
        dev = lay.flat_p.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(L.hav_adam_flat(C.c_void_p(lay.flat_p.data_ptr()), C.c_void_p(lay.flat_g.data_ptr()), C.c_void_p(self.m.data_ptr()),
                                       C.c_void_p(self.v.data_ptr()), lay.numel, C.c_void_p(self.state.data_ptr()), self.betas[0],
                                       self.betas[1], self.eps, float(grad_scale), 1, C.c_void_p(st)), "hav_adam_flat")
        styleunet.invalidate_caches()      # parameters changed behind torch's version counters

    def zero_grad(self):
        self.layout.flat_g.zero_()

    # ---- checkpoint compatibility with torch.optim.Adam (train_avatar.py:91,306 save / restore optimizer_state_dict)
    @property
    def param_groups(self):
        """One group, like the reference's optimisers; 'lr' reads the device-side learning rate."""
        return [{"params": list(self.layout.params), "lr": float(self.state[1]), "betas": self.betas, "eps": self.eps,
                 "weight_decay": 0, "amsgrad": False}]

    def state_dict(self):
        """torch.optim.Adam.state_dict() layout: per-parameter 'step' / 'exp_avg' / 'exp_avg_sq' sliced out of the flat moments."""
        lay = self.layout
        step = self.state[0].detach().clone().cpu()
        st = {}
        for i, (p, o) in enumerate(zip(lay.params, lay.offsets)):
            st[i] = {"step": step.clone(), "exp_avg": self.m[o:o + p.numel()].view(p.shape).clone(),
                     "exp_avg_sq": self.v[o:o + p.numel()].view(p.shape).clone()}
        group = {"lr": float(self.state[1]), "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(lay.params)))}
        return {"state": st, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts a torch.optim.Adam state_dict (or one written by state_dict() above) over the same parameter list: the
        moments are copied into the flat buffers and the step count / learning rate restored, so bias correction resumes where
        the checkpoint left off.  Parameters without saved state (never stepped) keep zero moments."""
        lay = self.layout
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(lay.params):
            raise ValueError("optimizer state has %d parameters, this FlatAdam has %d" % (len(ids), len(lay.params)))
        steps = set()
        with torch.no_grad():
            self.m.zero_(), self.v.zero_()
            for pid, p, o in zip(ids, lay.params, lay.offsets):
                st = sd["state"].get(pid)
                if st is None:
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError("optimizer state shape %s does not match parameter %s" % (tuple(st["exp_avg"].shape), tuple(p.shape)))
                self.m[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
                self.v[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(float(st["step"]))
            if len(steps) > 1:
                raise ValueError("FlatAdam keeps one step count; the state dict has %s" % sorted(steps))
            self.state[0:1].fill_(steps.pop() if steps else 0.0)
            self.state[1:2].fill_(float(groups[0]["lr"]))
        g = groups[0]
        self.betas, self.eps = (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"])


class _Bucket:
    __slots__ = ("flat", "params", "pending", "ready", "issued", "work")

    def __init__(self, flat, params):
        self.flat, self.params = flat, params
        self.pending, self.ready, self.issued, self.work = len(params), False, False, None


class GradSync:
    """Bucketed, backward-overlapped gradient all-reduce over `params` (an iterable of nn.Parameter).

        sync = GradSync(model.parameters())
        loss.backward()          # hooks issue the bucket all-reduces while backward is still running
        sync.finish()            # every .grad now holds the mean over ranks
        optimizer.step(); sync.zero_grad()

    `.grad` of every parameter is a view into its bucket and must stay one: use `sync.zero_grad()` (one memset per bucket)
    instead of `optimizer.zero_grad()` (whose default set_to_none=True would drop the views)."""

    def __init__(self, params, group=None, bucket_bytes=None, average=True, layout=None):
        """layout: a FlatLayout that already owns the gradient views (buckets become contiguous slices of its flat gradient
        buffer); None allocates per-bucket buffers here."""
        if bucket_bytes is None:      # 64 MiB unless HAV_GRAD_BUCKET_MB says otherwise (a tuning aid)
            import os

            bucket_bytes = int(os.environ.get("HAV_GRAD_BUCKET_MB", "64")) << 20
        self.group = group
        self.layout = layout
        if layout is not None:
            params = layout.params
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.average = average
        params = [p for p in params if p.requires_grad]
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        backend = dist.get_backend(group) if self.world > 1 else None
        self._avg_in_collective = bool(average and backend == "nccl")
        self.buckets, self._bucket_of, self._hooks = [], {}, []
        cur, cur_bytes, key = [], 0, None
        for p in reversed(self.params):
            k = (p.dtype, p.device)
            nbytes = p.numel() * p.element_size()
            if cur and (k != key or cur_bytes + nbytes > bucket_bytes):
                self._close(cur)
                cur, cur_bytes = [], 0
            key = k
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._next = 0
        self.issued_in_backward = 0   # buckets whose exchange started from inside backward (diagnostic)
        self.collectives = 0          # all-reduces issued since construction (bench / tests read it)
        self.bytes_per_step = sum(b.flat.numel() * b.flat.element_size() for b in self.buckets)

    def _close(self, params):
        if self.layout is not None:
            # `params` are consecutive parameters in REVERSE layout order: their slots form one contiguous slice
            idx = {id(p): i for i, p in enumerate(self.layout.params)}
            first, last = idx[id(params[-1])], idx[id(params[0])]
            lo = self.layout.offsets[first]
            hi = self.layout.offsets[last] + -(-self.layout.params[last].numel() // 32) * 32
            flat = self.layout.flat_g[lo:hi]
        else:
            total = sum(-(-p.numel() // 32) * 32 for p in params)        # 128-byte aligned slots
            flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        b = _Bucket(flat, params)
        off = 0
        for p in params:
            if self.layout is None:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += -(-p.numel() // 32) * 32
            self._bucket_of[id(p)] = len(self.buckets)
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.buckets.append(b)

    # ---- backward-time path
    def _on_grad(self, p):
        b = self.buckets[self._bucket_of[id(p)]]
        b.pending -= 1
        if b.pending == 0:
            b.ready = True
            self._issue_ready()

    def _issue(self, b):
        if self.world > 1:
            comm = None
            if b.flat.is_cuda:
                # backward may run on more than one stream (pipeline.two_stream_planes, run_parallel, ...): the collective must
                # follow everything enqueued so far on ANY of them.  It is issued from a dedicated staging stream that waits on
                # all of them, so that no compute stream is held up by a sibling's unfinished work (a wait on the issuing
                # stream itself would serialise the branches at every bucket boundary).
                from .pipeline import aux_stream, forked_streams

                dev = b.flat.device
                cur = torch.cuda.current_stream(dev)
                comm = aux_stream(dev, 15)
                capturing = torch.cuda.is_current_stream_capturing()
                comm.wait_stream(cur)
                for s in forked_streams(dev):
                    if s == cur or s == comm:
                        continue
                    if capturing:      # only streams that are part of this capture may be waited on (the others are idle)
                        with torch.cuda.stream(s):
                            if not torch.cuda.is_current_stream_capturing():
                                continue
                    comm.wait_stream(s)
                self._comm = comm
            op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
            if comm is not None:
                with torch.cuda.stream(comm):
                    b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
            else:
                b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
            self.collectives += 1
        b.ready = True

    def _issue_ready(self):
        """Issue every bucket that is complete and not yet on its way, in the order they became complete.  Every rank runs the
        same autograd graph through the same single-threaded engine, so that order is the same on every rank (and under a CUDA-
        graph capture it is frozen at capture time).  A strict bucket-index order would be safe too, but one straggler in an
        early bucket (a style MLP whose gradient is the last of its network to arrive sits in the first bucket of the reversed
        parameter list) then holds every later bucket back until the end of backward and the whole exchange is exposed."""
        for b in self.buckets:
            if b.ready and not b.issued:
                b.issued = True
                self._issue(b)
                self.issued_in_backward += 1

    def finish(self):
        """Issue the buckets backward did not complete, wait for all of them, apply the 1/world scale where the collective
        did not, and re-arm for the next step."""
        for b in self.buckets:
            if not b.issued:
                b.issued = True
                self._issue(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                b.work = None
            if self.world > 1 and self.average and not self._avg_in_collective:
                b.flat.mul_(1.0 / self.world)
        comm = getattr(self, "_comm", None)
        if comm is not None:           # the staging stream rejoins the caller's stream (a capture must end with every fork joined)
            torch.cuda.current_stream(comm.device).wait_stream(comm)
        for b in self.buckets:
            b.pending, b.ready, b.issued = len(b.params), False, False
            for p in b.params:       # an optimiser's zero_grad(set_to_none=True) or a grad replaced by autograd breaks the views
                if p.grad is None or p.grad.data_ptr() < b.flat.data_ptr() or \
                        p.grad.data_ptr() >= b.flat.data_ptr() + b.flat.numel() * b.flat.element_size():
                    raise RuntimeError("GradSync: a parameter's .grad no longer aliases its bucket; use GradSync.zero_grad()")
        self._next = 0

    def zero_grad(self):
        for b in self.buckets:
            b.flat.zero_()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
