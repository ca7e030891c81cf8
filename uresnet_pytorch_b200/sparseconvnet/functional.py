"""torch.autograd.Functions over the C-ABI (include/b200scn.h).

Each Function is the forward/dgrad/wgrad triple of one scn layer type used by
/root/reference/uresnet/models/uresnet_sparse.py:19-24.  They only pass raw device pointers,
row strides and the current stream; all arithmetic happens in libb200scn.
"""
import torch

from .. import _lib
from . import config


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rowmajor(x):
    return x if (x.dim() == 2 and x.stride(1) == 1 and x.stride(0) >= x.shape[1]) else x.contiguous()


def _check_cuda_f32(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b200 sparseconvnet: tensors must be on a CUDA device (there is no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError("b200 sparseconvnet: features and weights are float32 (got %s)" % t.dtype)


_side_streams = {}
_side_dirty = set()
_side_keepalive = []  # buffers a side-stream kernel still reads/writes; released at the join


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def join_side_streams():
    """Make the current stream wait for every weight-gradient kernel issued on a side stream."""
    for key in list(_side_dirty):
        torch.cuda.current_stream(torch.device(*key)).wait_stream(_side_streams[key])
    _side_dirty.clear()
    # frees happen on the main stream *after* it waited for the side stream: stream-ordered reuse is safe
    del _side_keepalive[:]


class GradSink(object):
    """One flat fp32 gradient buffer with a view per parameter (the buffer the per-step NCCL
    all-reduce runs over, replacing /root/reference/uresnet/ops.py:8-60's per-replica reduce-add).

    While a sink is ACTIVE (`with sink.active():`) the backward kernels of this package write -- always
    accumulating -- straight into the parameter's view and hand autograd no gradient tensor at all; `p.grad`
    IS that view, permanently.  `zero()` (one memset of the flat buffer) starts a new accumulation, so several
    forwards can be accumulated before one optimiser step (R/uresnet/trainval.py:21-30,59-69), a weight used
    twice sums correctly, and parameters whose gradient arrives through plain torch autograd (an unfused
    torch.nn.Linear head) accumulate into the same memory.  Because nothing consumes the views before the owner
    has joined the side streams, weight-gradient kernels may run on a side stream without racing autograd."""

    def __init__(self, params):
        self.params = [p for p in params]
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # every view starts 16-byte aligned
        self.flat = torch.zeros(max(total, 4), dtype=torch.float32, device=dev)
        self.numel = total
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
        self.offsets = offs
        self._on = False
        for p, v in zip(self.params, self.views):
            p._b200_sink = (self, v)
            p.grad = v

    def zero(self):
        self.flat.zero_()

    def active(self):
        return _SinkScope(self)


class _SinkScope(object):
    def __init__(self, sink):
        self.sink = sink

    def __enter__(self):
        self.prev = self.sink._on
        self.sink._on = True
        for p, v in zip(self.sink.params, self.sink.views):
            if p.grad is not v:     # zero_grad(set_to_none=True) or an optimiser replaced it
                p.grad = v
        return self.sink

    def __exit__(self, *exc):
        self.sink._on = self.prev


def _sink_view(param):
    """The active sink's view for `param`, or None (-> the gradient goes back through autograd)."""
    ent = getattr(param, "_b200_sink", None) if param is not None else None
    if ent is not None and ent[0]._on:
        return ent[1]
    return None


def _wgrad(A, B, table, K, gather_b, rows, n_dev, w_shape, math, gpl=None, subm=False, sink=None, rows_hint=0):
    """dW [K, Ca, Cb].  `gpl`: the layer's FORWARD pair list (its out rows index B) -- when given, and the
    shapes allow, the weight gradient walks it (csrc/wgrad_gpl.cu) instead of re-compacting the table.
    `sink`: accumulate into this preallocated view (GradSink) and return None."""
    Ca, Cb = A.shape[1], B.shape[1]
    L = _lib.load()
    aligned = (Ca % 16 == 0 and Cb % 16 == 0 and A.stride(0) % 4 == 0 and B.stride(0) % 4 == 0
               and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0)
    # tcgen05 kernel (sparse_tc.cu): rule tables that gather on the A side (subm, strided conv, NiN)
    use_tc = (config.use_tc and config.use_tc_wgrad and math != 0 and aligned and gather_b == 0 and K <= 27
              and (table is not None or K == 1) and min(Ca, Cb) >= max(config.tc_wgrad_min_channels, 16))
    use_gpl = (not use_tc and gpl is not None and gpl.get("koff") is not None and config.use_gpl and config.wgrad_gpl
               and math != 0 and aligned)
    hint = rows_hint if 0 < rows_hint < rows else rows
    if use_tc:
        nbytes = L.b200scn_wgrad_tc_scratch_bytes(hint, K, Ca, Cb)
    elif use_gpl:
        nbytes = L.b200scn_wgrad_gpl_scratch_bytes(K, Ca, Cb)
    else:
        nbytes = L.b200scn_wgrad_scratch_bytes(rows, K, Ca, Cb)

    # every buffer is allocated on the main stream; only the kernels move to the side stream
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    dW = sink if sink is not None else torch.empty(w_shape, dtype=torch.float32, device=A.device)
    acc = 1 if sink is not None else 0

    def run():
        if use_tc:
            _lib.call("b200scn_wgrad_tc", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), _lib.ptr(table),
                      table.stride(0) if table is not None else 0, K, rows, n_dev, hint, Ca, Cb, dW.data_ptr(), acc,
                      math, scratch.data_ptr(), nbytes, _stream())
            return
        if use_gpl:
            _lib.call("b200scn_wgrad_gpl", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), gpl["tile_rows"],
                      gpl["g_rec"].data_ptr(), gpl["koff"].data_ptr(), K, int(subm), rows, n_dev,
                      gpl.get("rows_hint", 0), Ca, Cb, dW.data_ptr(), acc, math, scratch.data_ptr(), nbytes, _stream())
            return
        _lib.call("b200scn_wgrad_table", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), _lib.ptr(table),
                  table.stride(0) if table is not None else 0, K, gather_b, rows, n_dev, Ca, Cb, dW.data_ptr(), acc,
                  math, scratch.data_ptr(), nbytes, _stream())

    # Side-stream overlap only with a sink: a gradient tensor handed to autograd is consumed (AccumulateGrad,
    # hooks, a second backward) on the main stream before anybody joins the side stream.
    if not (config.overlap_wgrad and sink is not None):
        run()
        return None if sink is not None else dW
    main = torch.cuda.current_stream(A.device)
    side = _side_stream(A.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        run()
    _side_keepalive.append((scratch, A, B, table, gpl, dW))
    _side_dirty.add((A.device.type, A.device.index))
    return None


def _gpl_ok(gpl, x, Cin, Cout, math):
    return (gpl is not None and config.use_gpl and math != 0 and Cin % 16 == 0 and Cout % 16 == 0
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0)


# Fragment-ordered weight copies (b200scn_pack_weights), one per (weight, direction).  The copies live ON the
# parameter object (`_b200_pack`), so they are freed with it and two models never see each other's entries.  A
# copy is refreshed when the weight's version counter moved; whoever updates weights behind autograd's back
# (FusedAdam writes through raw pointers, inside a CUDA graph) calls `repack_params` right after, which refreshes
# every copy of the given parameters with ONE launch over a job table it returns -- the caller keeps that table
# alive as long as a captured graph may replay the launch.
def _pack_entry(W, W3, Cin, Cout, w_si, w_so, kflip, kind="mma"):
    """kind 'mma': fragment order of the mma.sync pair-list kernels (conv_gpl.cu); 'tc': K-major 32-channel blocks,
    high and low tf32 parts, of the tcgen05 kernel (sparse_tc.cu)."""
    cache = getattr(W, "_b200_pack", None)
    if cache is None:
        cache = W._b200_pack = {}
    key = (kind, w_so == 1, kflip)
    ent = cache.get(key)
    if ent is None or ent["ptr"] != W3.data_ptr() or ent["shape"] != tuple(W3.shape):
        K = W3.shape[0]
        if kind == "tc":
            elems = int(_lib.load().b200scn_conv_tc_packed_elems(K, Cin, Cout))
            buf = torch.empty(2 * elems, dtype=torch.float32, device=W3.device)  # [high | low]
            job = torch.tensor([W3.data_ptr(), buf.data_ptr(), buf.data_ptr() + 4 * elems, K, Cin, Cout,
                                W3.shape[1] * W3.shape[2], w_si, w_so, kflip, 0, 0], dtype=torch.int64).to(W3.device)
        else:
            elems = K * Cin * Cout
            buf = torch.empty(elems, dtype=torch.float32, device=W3.device)
            job = torch.tensor([W3.data_ptr(), buf.data_ptr(), K, Cin, Cout, W3.shape[1] * W3.shape[2], w_si, w_so,
                                kflip, 0], dtype=torch.int64).to(W3.device)
        ent = dict(ptr=W3.data_ptr(), shape=tuple(W3.shape), buf=buf, job=job, version=None, elems=elems, kind=kind)
        cache[key] = ent
    return ent


_PACK_FN = {"mma": "b200scn_pack_weights", "tc": "b200scn_pack_weights_tc"}


def packed_weights(W, W3, Cin, Cout, w_si, w_so, kflip, kind="mma"):
    """Cache entry of the packed copy of W3 (a view of parameter W) for one direction, refreshed if stale."""
    ent = _pack_entry(W, W3, Cin, Cout, w_si, w_so, kflip, kind)
    if ent["version"] != W._version:
        _lib.call(_PACK_FN[kind], ent["job"].data_ptr(), 1, ent["elems"], _stream())
        ent["version"] = W._version
    return ent


def repack_params(params, jobs=None):
    """Refresh every cached copy of `params` with ONE launch per kind.  Returns the job tables (pass them back
    in to reuse them; they are rebuilt only when the set of copies changed)."""
    ents = [e for p in params for e in getattr(p, "_b200_pack", {}).values()]
    if not ents:
        return jobs
    ids = tuple(id(e) for e in ents)
    if jobs is None or jobs[0] != ids:
        tabs = {}
        for kind in _PACK_FN:
            es = [e for e in ents if e["kind"] == kind]
            if es:
                tabs[kind] = (torch.cat([e["job"] for e in es]), len(es), max(e["elems"] for e in es))
        jobs = (ids, tabs, ents)
    for kind, (tab, n, mx) in jobs[1].items():
        _lib.call(_PACK_FN[kind], tab.data_ptr(), n, mx, _stream())
    for p in params:
        for e in getattr(p, "_b200_pack", {}).values():
            e["version"] = p._version
    return jobs


def _tc_ok(x, K, Cin, Cout, math):
    """The tcgen05 rule-table kernel (csrc/sparse_tc.cu) takes this layer call.  It wins where the layer is wide
    (config.tc_min_channels: measured cross-over, profiles/README.md round 2): each 128-row operand tile is read
    by the tensor core once per 8 contraction elements whatever N is, so narrow layers are shared-memory bound."""
    return (config.use_tc and math != 0 and K <= 27 and Cin % 16 == 0 and Cout % 16 == 0
            and min(Cin, Cout) >= max(config.tc_min_channels, 16)
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0)


def _conv_tc(x, W, W3, w_si, w_so, kflip, table, n_out, n_out_dev, Cin, Cout, math, rows_hint=0, res=None):
    """out[o] = (res[o] +) sum_k x[table[k][o]] @ W[k'] on the tcgen05 tensor cores (table None: identity, K = 1)."""
    K = W3.shape[0]
    ent = packed_weights(W, W3, Cin, Cout, w_si, w_so, kflip, "tc")
    out = torch.empty((n_out, Cout), dtype=torch.float32, device=x.device)
    hi = ent["buf"].data_ptr()
    _lib.call("b200scn_conv_tc", x.data_ptr(), x.stride(0), hi, hi + 4 * ent["elems"], _lib.ptr(table),
              table.stride(0) if table is not None else 0, K, n_out, n_out_dev, rows_hint, Cin, Cout, out.data_ptr(),
              out.stride(0), _lib.ptr(res), res.stride(0) if res is not None else 0, math, _stream())
    return out


def _conv_gpl(x, W, W3, w_si, w_so, kflip, gpl, n_out, n_out_dev, Cin, Cout, math, res=None):
    """out[o] = sum over the pair list's (in -> o, k) rules of x[in] @ W[k] (strided view of W3)."""
    K = W3.shape[0]
    out = torch.empty((n_out, Cout), dtype=torch.float32, device=x.device)
    if gpl.get("variant") is not None:
        # submanifold list: the large-input kernel wants the dense offsets first (metadata.subm_gpl)
        want = _lib.load().b200scn_conv_gpl_layout(gpl["tile_rows"], int(gpl["tile_end"] is None), n_out, Cin, Cout, K, 1,
                                                   math)
        if want:
            gpl = gpl["variant"](want)
    Wp = packed_weights(W, W3, Cin, Cout, w_si, w_so, kflip)["buf"].data_ptr() if config.pack_weights else None
    _lib.call("b200scn_conv_gpl", x.data_ptr(), x.stride(0), W3.data_ptr(), Wp, W3.shape[1] * W3.shape[2], w_si, w_so,
              kflip, gpl["tile_rows"], gpl.get("layout", 0), gpl["tile_off"].data_ptr(), _lib.ptr(gpl["tile_end"]),
              gpl["g_rec"].data_ptr(), K,
              n_out, n_out_dev, gpl.get("rows_hint", 0),
              Cin, Cout, out.data_ptr(), out.stride(0), 0, _lib.ptr(res), res.stride(0) if res is not None else 0, math,
              _stream())
    return out


class TableConv(torch.autograd.Function):
    """out[o] = sum_k x[table[k][o]] @ W[k]; table rows index x.  Used by SubmanifoldConvolution
    (table = neighbour table, symmetric), Convolution (table = child table) and
    NetworkInNetwork (table = None)."""

    @staticmethod
    def forward(ctx, x, W, table, n_out, n_out_dev, n_in, n_in_dev, symmetric, parent, kofs, gpl_fwd=None,
                gpl_bwd=None, rows_hint=0, res=None):
        """res: optional [n_out, Cout] tensor added to the result (the shortcut of a residual block, fused into
        the convolution's epilogue on the tcgen05 path)."""
        _check_cuda_f32(x, W)
        x = _rowmajor(x)
        W3 = W if W.dim() == 3 else W.unsqueeze(0)
        W3 = W3.contiguous()
        K, Cin, Cout = W3.shape
        math = config.math_mode
        hint = gpl_fwd.get("rows_hint", 0) if gpl_fwd is not None else rows_hint
        if res is not None:
            res = _rowmajor(res)
        fused_res = False
        if _tc_ok(x, K, Cin, Cout, math):
            fused_res = res is not None and res.stride(0) % 4 == 0 and res.data_ptr() % 16 == 0
            out = _conv_tc(x, W, W3, Cout, 1, 0, table, n_out, n_out_dev, Cin, Cout, math, hint,
                           res if fused_res else None)
        elif _gpl_ok(gpl_fwd, x, Cin, Cout, math):
            fused_res = res is not None
            out = _conv_gpl(x, W, W3, Cout, 1, 0, gpl_fwd, n_out, n_out_dev, Cin, Cout, math, res)
        else:
            out = torch.empty((n_out, Cout), dtype=torch.float32, device=x.device)
            _lib.call("b200scn_conv_table", x.data_ptr(), x.stride(0), W3.data_ptr(), Cin * Cout, Cout, 1, 0,
                      _lib.ptr(table), table.stride(0) if table is not None else 0, K, n_out, n_out_dev, Cin, Cout,
                      out.data_ptr(), out.stride(0), 0, math, _stream())
        if res is not None and not fused_res:
            out = out + res[:n_out]
        ctx.save_for_backward(x, W3)
        ctx.misc = (table, n_out, n_out_dev, n_in, n_in_dev, symmetric, parent, kofs, W.shape, math, gpl_bwd)
        ctx.gpl_fwd = gpl_fwd
        ctx.w_param = W
        ctx.rows_hint = hint
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W3 = ctx.saved_tensors
        table, n_out, n_out_dev, n_in, n_in_dev, symmetric, parent, kofs, w_shape, math, gpl_bwd = ctx.misc
        dout = _rowmajor(dout)
        K, Cin, Cout = W3.shape
        dx = dW = None
        if ctx.needs_input_grad[0] and (table is None or symmetric) and _tc_ok(dout, K, Cout, Cin, math):
            # dx[i] = sum_k dout[table[k][i]] @ W[K-1-k]^T   (the neighbour table is symmetric)
            dx = _conv_tc(dout, ctx.w_param, W3, 1, Cout, 1, table, n_in, n_in_dev, Cout, Cin, math, ctx.rows_hint)
        elif ctx.needs_input_grad[0] and _gpl_ok(gpl_bwd, dout, Cout, Cin, math):
            # same pair structure, transposed weights; the symmetric (submanifold) list needs the offset flip
            dx = _conv_gpl(dout, ctx.w_param, W3, 1, Cout, 1 if symmetric else 0, gpl_bwd, n_in, n_in_dev, Cout, Cin,
                           math)
        elif ctx.needs_input_grad[0]:
            dx = torch.empty((n_in, Cin), dtype=torch.float32, device=x.device)
            if table is None or symmetric:
                # dx[i] = sum_k dout[table[k][i]] @ W[K-1-k]^T   (neighbour table is symmetric)
                _lib.call("b200scn_conv_table", dout.data_ptr(), dout.stride(0), W3.data_ptr(), Cin * Cout, 1, Cout,
                          1, _lib.ptr(table), table.stride(0) if table is not None else 0, K, n_in, n_in_dev, Cout,
                          Cin, dx.data_ptr(), dx.stride(0), 0, math, _stream())
            else:
                # strided conv: every fine row has exactly one parent
                _lib.call("b200scn_conv_parent", dout.data_ptr(), dout.stride(0), W3.data_ptr(), Cin * Cout, 1, Cout,
                          parent.data_ptr(), kofs.data_ptr(), K, n_in, n_in_dev, Cout, Cin, dx.data_ptr(),
                          dx.stride(0), math, _stream())
        if ctx.needs_input_grad[1]:
            dW = _wgrad(x, dout, table, K, 0, n_out, n_out_dev, w_shape, math, ctx.gpl_fwd, symmetric,
                        sink=_sink_view(ctx.w_param), rows_hint=ctx.rows_hint)
        d_res = dout if (len(ctx.needs_input_grad) > 13 and ctx.needs_input_grad[13]) else None
        return dx, dW, None, None, None, None, None, None, None, None, None, None, None, d_res


class ParentConv(torch.autograd.Function):
    """out[p] = x[parent[p]] @ W[kofs[p]] (Deconvolution with filter == stride)."""

    @staticmethod
    def forward(ctx, x, W, parent, kofs, child, n_fine, n_fine_dev, n_coarse, n_coarse_dev, gpl_fwd=None,
                gpl_bwd=None):
        _check_cuda_f32(x, W)
        x = _rowmajor(x)
        W3 = W.contiguous()
        K, Cin, Cout = W3.shape
        math = config.math_mode
        if _gpl_ok(gpl_fwd, x, Cin, Cout, math):
            out = _conv_gpl(x, W, W3, Cout, 1, 0, gpl_fwd, n_fine, n_fine_dev, Cin, Cout, math)
        else:
            out = torch.empty((n_fine, Cout), dtype=torch.float32, device=x.device)
            _lib.call("b200scn_conv_parent", x.data_ptr(), x.stride(0), W3.data_ptr(), Cin * Cout, Cout, 1,
                      parent.data_ptr(), kofs.data_ptr(), K, n_fine, n_fine_dev, Cin, Cout, out.data_ptr(),
                      out.stride(0), math, _stream())
        ctx.save_for_backward(x, W3)
        ctx.misc = (parent, kofs, child, n_fine, n_fine_dev, n_coarse, n_coarse_dev, math, gpl_bwd)
        ctx.gpl_fwd = gpl_fwd
        ctx.w_param = W
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W3 = ctx.saved_tensors
        parent, kofs, child, n_fine, n_fine_dev, n_coarse, n_coarse_dev, math, gpl_bwd = ctx.misc
        dout = _rowmajor(dout)
        K, Cin, Cout = W3.shape
        dx = dW = None
        if ctx.needs_input_grad[0] and _gpl_ok(gpl_bwd, dout, Cout, Cin, math):
            dx = _conv_gpl(dout, ctx.w_param, W3, 1, Cout, 0, gpl_bwd, n_coarse, n_coarse_dev, Cout, Cin, math)
        elif ctx.needs_input_grad[0]:
            # dx[j] = sum_k dout[child[k][j]] @ W[k]^T
            dx = torch.empty((n_coarse, Cin), dtype=torch.float32, device=x.device)
            _lib.call("b200scn_conv_table", dout.data_ptr(), dout.stride(0), W3.data_ptr(), Cin * Cout, 1, Cout, 0,
                      child.data_ptr(), child.stride(0), K, n_coarse, n_coarse_dev, Cout, Cin, dx.data_ptr(),
                      dx.stride(0), 0, math, _stream())
        if ctx.needs_input_grad[1]:
            # dW[k] = sum_j x[j]^T @ dout[child[k][j]]
            if ctx.gpl_fwd is not None and ctx.gpl_fwd.get("koff") is not None:
                # the forward list runs over the FINE rows (in = parent row, out = fine row): rows = n_fine
                dW = _wgrad(x, dout, child, K, 1, n_fine, n_fine_dev, W3.shape, math, ctx.gpl_fwd, False,
                            sink=_sink_view(ctx.w_param))
            else:
                dW = _wgrad(x, dout, child, K, 1, n_coarse, n_coarse_dev, W3.shape, math,
                            sink=_sink_view(ctx.w_param))
        return dx, dW, None, None, None, None, None, None, None, None, None


class BatchNormAct(torch.autograd.Function):
    """Fused BatchNorm + leaky ReLU (leak = 0: ReLU, leak = 1: none)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, retention, leak, training, rows, n_dev,
                sync=None, rows_hint=0):
        """sync: the module's two-word barrier state (zeroed once) -> single-launch kernels for small inputs."""
        _check_cuda_f32(x, weight, bias)
        x = _rowmajor(x)
        C = x.shape[1]
        L = _lib.load()
        y = torch.empty((rows, C), dtype=torch.float32, device=x.device)
        if training:
            nbytes = L.b200scn_bn_scratch_bytes(rows, C)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            mean = torch.empty(C, dtype=torch.float32, device=x.device)
            invstd = torch.empty(C, dtype=torch.float32, device=x.device)
            _lib.call("b200scn_bn_fwd_train_sync", x.data_ptr(), x.stride(0), rows, n_dev, rows_hint, C,
                      _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(running_mean), _lib.ptr(running_var), eps, retention,
                      leak, y.data_ptr(), y.stride(0), mean.data_ptr(), invstd.data_ptr(), _lib.ptr(sync),
                      scratch.data_ptr(), nbytes, _stream())
        else:
            mean = running_mean
            invstd = torch.rsqrt(running_var + eps)
            _lib.call("b200scn_bn_fwd_eval", x.data_ptr(), x.stride(0), rows, n_dev, C, _lib.ptr(weight),
                      _lib.ptr(bias), running_mean.data_ptr(), running_var.data_ptr(), eps, leak, y.data_ptr(),
                      y.stride(0), _stream())
        ctx.save_for_backward(x, weight, bias, mean, invstd)
        ctx.misc = (leak, training, rows, n_dev, sync, rows_hint)
        ctx.params = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, invstd = ctx.saved_tensors
        leak, training, rows, n_dev, sync, rows_hint = ctx.misc
        dy = _rowmajor(dy)
        C = x.shape[1]
        L = _lib.load()
        nbytes = L.b200scn_bn_scratch_bytes(rows, C)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        dx = torch.empty((rows, C), dtype=torch.float32, device=x.device)
        # affine gradients: straight into the trainer's flat gradient buffer when one is active
        sg, sb = _sink_view(ctx.params[0]), _sink_view(ctx.params[1])
        sunk = sg is not None and sb is not None
        dgamma = sg if sunk else torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = sb if sunk else torch.empty(C, dtype=torch.float32, device=x.device)
        _lib.call("b200scn_bn_bwd_sync", x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), rows, n_dev,
                  rows_hint, C, _lib.ptr(weight), _lib.ptr(bias), mean.data_ptr(), invstd.data_ptr(), leak,
                  int(training) | (2 if sunk else 0), dx.data_ptr(), dx.stride(0), dgamma.data_ptr(),
                  dbeta.data_ptr(), _lib.ptr(sync), scratch.data_ptr(), nbytes, _stream())
        if sunk:
            dgamma = dbeta = None
        return (dx, dgamma if weight is not None else None, dbeta if bias is not None else None, None, None, None,
                None, None, None, None, None, None, None)


class InputGather(torch.autograd.Function):
    """InputLayer mode 3: active row = sum of its input points, in input order."""

    @staticmethod
    def forward(ctx, feats, md, rows, n_dev):
        _check_cuda_f32(feats)
        feats = _rowmajor(feats)
        C = feats.shape[1]
        out = torch.empty((rows, C), dtype=torch.float32, device=feats.device)
        _lib.call("b200scn_segsum_rows", feats.data_ptr(), feats.stride(0), md.first_idx.data_ptr(),
                  md.seg_off.data_ptr(), md.seg.data_ptr(), rows, n_dev, C, out.data_ptr(), out.stride(0), _stream())
        ctx.md = md
        return out

    @staticmethod
    def backward(ctx, dout):
        md = ctx.md
        dout = _rowmajor(dout)
        C = dout.shape[1]
        din = torch.empty((md.n_in, C), dtype=torch.float32, device=dout.device)
        _lib.call("b200scn_gather_rows", dout.data_ptr(), dout.stride(0), md.in2row.data_ptr(), md.n_in,
                  _lib.ptr(md.n_in_dev), C, din.data_ptr(), din.stride(0), _stream())
        return din, None, None, None


class OutputGather(torch.autograd.Function):
    """OutputLayer: every original input point receives its active row."""

    @staticmethod
    def forward(ctx, feats, md, rows, n_dev):
        _check_cuda_f32(feats)
        feats = _rowmajor(feats)
        C = feats.shape[1]
        # padded rows (static regime) must hold finite values: the head and the loss read them
        alloc = torch.zeros if md.n_in_dev is not None else torch.empty
        out = alloc((md.n_in, C), dtype=torch.float32, device=feats.device)
        _lib.call("b200scn_gather_rows", feats.data_ptr(), feats.stride(0), md.in2row.data_ptr(), md.n_in,
                  _lib.ptr(md.n_in_dev), C, out.data_ptr(), out.stride(0), _stream())
        ctx.md, ctx.rows, ctx.n_dev = md, rows, n_dev
        return out

    @staticmethod
    def backward(ctx, dout):
        md = ctx.md
        dout = _rowmajor(dout)
        C = dout.shape[1]
        d = torch.empty((ctx.rows, C), dtype=torch.float32, device=dout.device)
        _lib.call("b200scn_segsum_rows", dout.data_ptr(), dout.stride(0), md.first_idx.data_ptr(),
                  md.seg_off.data_ptr(), md.seg.data_ptr(), ctx.rows, ctx.n_dev, C, d.data_ptr(), d.stride(0),
                  _stream())
        return d, None, None, None


class HeadLoss(torch.autograd.Function):
    """Fused Linear(m -> nc) + per-event mean cross-entropy + per-event accuracy (csrc/headloss.cu;
    reference: uresnet_sparse.py:25,36,44-82).  Returns a tensor [4] = (total_loss, total_acc, rows with a
    sample id outside [0, nb), rows with a label outside [0, nc)).  The backward pass was already computed by
    the forward kernel for an upstream gradient of 1 on total_loss; `backward` scales it by the incoming
    gradient of element 0."""

    @staticmethod
    def forward(ctx, feat, W, b, label, ids, nb, pix_w, rows, n_dev, unit_grad):
        _check_cuda_f32(feat, W, b)
        feat = _rowmajor(feat)
        Wc = W.contiguous()
        label = label.reshape(-1)
        if label.dtype != torch.float32:
            label = label.float()
        if ids.dtype not in (torch.int32, torch.int64, torch.float32, torch.float64):
            ids = ids.long()
        codes = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}
        m, nc = feat.shape[1], Wc.shape[0]
        L = _lib.load()
        nbytes = L.b200scn_head_loss_scratch_bytes(rows, m, nc, nb)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=feat.device)
        out = torch.empty(4, dtype=torch.float32, device=feat.device)
        need = feat.requires_grad or W.requires_grad
        dfeat = torch.empty((rows, m), dtype=torch.float32, device=feat.device) if need else None
        # with an upstream gradient of exactly 1 the parameter gradients can go straight into the flat buffer
        sW, sb = (_sink_view(W), _sink_view(b)) if (need and unit_grad and b is not None and Wc is W) else (None, None)
        sunk = sW is not None and sb is not None
        dW = (sW if sunk else torch.empty_like(Wc)) if need else None
        db = (sb if sunk else torch.empty(nc, dtype=torch.float32, device=feat.device)) if need else None
        if pix_w is not None:
            pix_w = pix_w.reshape(-1).float()
        _lib.call("b200scn_head_loss", feat.data_ptr(), feat.stride(0), Wc.data_ptr(), _lib.ptr(b), label.data_ptr(),
                  label.stride(0), _lib.ptr(pix_w), pix_w.stride(0) if pix_w is not None else 0, ids.data_ptr(),
                  codes[ids.dtype], ids.stride(0), rows, n_dev, m, nc, nb, None, 0, out.data_ptr(), _lib.ptr(dfeat),
                  dfeat.stride(0) if dfeat is not None else 0, _lib.ptr(dW), _lib.ptr(db), 1 if sunk else 0,
                  scratch.data_ptr(), nbytes, _stream())
        ctx.grads = (dfeat, None, None) if sunk else (dfeat, dW, db)
        ctx.unit_grad = unit_grad
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        dfeat, dW, db = ctx.grads
        if not ctx.unit_grad:
            g = gout[0]
            dfeat = dfeat * g
            dW = dW * g if dW is not None else None
            db = db * g if db is not None else None
        return dfeat, dW, (db if ctx.has_bias else None), None, None, None, None, None, None, None


def head_loss(feat, linear, label, ids, nb, pix_w=None, rows=None, n_dev=None, unit_grad=False, full=False):
    """(total_loss, total_acc) of a torch.nn.Linear head + the reference's SegmentationLoss on `feat`.
    full=True returns the raw [4] tensor (loss, acc, bad sample ids, bad labels) instead."""
    rows = feat.shape[0] if rows is None else rows
    out = HeadLoss.apply(feat, linear.weight, linear.bias, label, ids, nb, pix_w, rows, n_dev, unit_grad)
    if full:
        return out
    return out[0], out[1].detach()
