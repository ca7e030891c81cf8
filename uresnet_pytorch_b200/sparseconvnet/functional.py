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


def _wgrad(A, B, table, K, gather_b, rows, n_dev, w_shape, math, gpl=None, subm=False):
    """dW [K, Ca, Cb].  `gpl`: the layer's FORWARD pair list (its out rows index B) -- when given, and the
    shapes allow, the weight gradient walks it (csrc/wgrad_gpl.cu) instead of re-compacting the table."""
    Ca, Cb = A.shape[1], B.shape[1]
    L = _lib.load()
    use_gpl = (gpl is not None and gpl.get("koff") is not None and config.use_gpl and config.wgrad_gpl and math != 0
               and Ca % 16 == 0 and Cb % 16 == 0 and A.stride(0) % 4 == 0 and B.stride(0) % 4 == 0
               and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0)
    nbytes = L.b200scn_wgrad_gpl_scratch_bytes(K, Ca, Cb) if use_gpl else L.b200scn_wgrad_scratch_bytes(rows, K, Ca, Cb)

    # every buffer is allocated on the main stream; only the kernels move to the side stream
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    dW = torch.empty(w_shape, dtype=torch.float32, device=A.device)

    def run():
        if use_gpl:
            _lib.call("b200scn_wgrad_gpl", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), gpl["tile_rows"],
                      gpl["g_rec"].data_ptr(), gpl["koff"].data_ptr(), K, int(subm), rows, n_dev,
                      gpl.get("rows_hint", 0), Ca, Cb, dW.data_ptr(), 0, math, scratch.data_ptr(), nbytes, _stream())
            return
        _lib.call("b200scn_wgrad_table", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), _lib.ptr(table),
                  table.stride(0) if table is not None else 0, K, gather_b, rows, n_dev, Ca, Cb, dW.data_ptr(), 0,
                  math, scratch.data_ptr(), nbytes, _stream())

    if not config.overlap_wgrad:
        run()
        return dW
    main = torch.cuda.current_stream(A.device)
    side = _side_stream(A.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        run()
    _side_keepalive.append((scratch, A, B, table, gpl))
    _side_dirty.add((A.device.type, A.device.index))
    return dW


def _gpl_ok(gpl, x, Cin, Cout, math):
    return (gpl is not None and config.use_gpl and math != 0 and Cin % 16 == 0 and Cout % 16 == 0
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0)


# Fragment-ordered weight copies (b200scn_pack_weights), one per (weight storage, direction).  An entry is
# re-packed when the weight's version counter moved; a trainer that steps the optimiser inside a CUDA graph
# calls repack_all() right after it, so the (captured) forward never finds a stale copy.
_pack_cache = {}
_pack_all_jobs = None


def _pack_entry(W3, Cin, Cout, w_si, w_so, kflip):
    key = (W3.data_ptr(), w_so == 1, kflip)
    ent = _pack_cache.get(key)
    if ent is None or ent["shape"] != tuple(W3.shape) or ent["device"] != W3.device:
        global _pack_all_jobs
        K = W3.shape[0]
        buf = torch.empty(K * Cin * Cout, dtype=torch.float32, device=W3.device)
        job = torch.tensor([W3.data_ptr(), buf.data_ptr(), K, Cin, Cout, W3.shape[1] * W3.shape[2], w_si, w_so, kflip,
                            0], dtype=torch.int64).to(W3.device)
        ent = dict(shape=tuple(W3.shape), device=W3.device, buf=buf, job=job, version=None, W=W3,
                   elems=K * Cin * Cout)
        _pack_cache[key] = ent
        _pack_all_jobs = None
    return ent


def packed_weights(W3, Cin, Cout, w_si, w_so, kflip):
    """Device pointer of the fragment-ordered copy of W3 for the direction (Cin, Cout, w_si, w_so, kflip)."""
    ent = _pack_entry(W3, Cin, Cout, w_si, w_so, kflip)
    if ent["version"] != W3._version:
        _lib.call("b200scn_pack_weights", ent["job"].data_ptr(), 1, ent["elems"], _stream())
        ent["version"] = W3._version
    return ent["buf"].data_ptr()


def repack_all():
    """Refresh every cached copy with ONE launch (after an optimiser step / a state-dict load)."""
    global _pack_all_jobs
    if not _pack_cache:
        return
    ents = list(_pack_cache.values())
    if _pack_all_jobs is None:
        _pack_all_jobs = torch.cat([e["job"] for e in ents])
    _lib.call("b200scn_pack_weights", _pack_all_jobs.data_ptr(), len(ents), max(e["elems"] for e in ents), _stream())
    for e in ents:
        e["version"] = e["W"]._version


def clear_pack_cache():
    global _pack_all_jobs
    _pack_cache.clear()
    _pack_all_jobs = None


def _conv_gpl(x, W3, w_si, w_so, kflip, gpl, n_out, n_out_dev, Cin, Cout, math):
    """out[o] = sum over the pair list's (in -> o, k) rules of x[in] @ W[k] (strided view of W3)."""
    K = W3.shape[0]
    out = torch.empty((n_out, Cout), dtype=torch.float32, device=x.device)
    Wp = packed_weights(W3, Cin, Cout, w_si, w_so, kflip) if config.pack_weights else None
    _lib.call("b200scn_conv_gpl", x.data_ptr(), x.stride(0), W3.data_ptr(), Wp, W3.shape[1] * W3.shape[2], w_si, w_so,
              kflip, gpl["tile_rows"], gpl["tile_off"].data_ptr(), _lib.ptr(gpl["tile_end"]), gpl["g_rec"].data_ptr(), K,
              n_out, n_out_dev, gpl.get("rows_hint", 0),
              Cin, Cout, out.data_ptr(), out.stride(0), 0, math, _stream())
    return out


class TableConv(torch.autograd.Function):
    """out[o] = sum_k x[table[k][o]] @ W[k]; table rows index x.  Used by SubmanifoldConvolution
    (table = neighbour table, symmetric), Convolution (table = child table) and
    NetworkInNetwork (table = None)."""

    @staticmethod
    def forward(ctx, x, W, table, n_out, n_out_dev, n_in, n_in_dev, symmetric, parent, kofs, gpl_fwd=None,
                gpl_bwd=None):
        _check_cuda_f32(x, W)
        x = _rowmajor(x)
        W3 = W if W.dim() == 3 else W.unsqueeze(0)
        W3 = W3.contiguous()
        K, Cin, Cout = W3.shape
        math = config.math_mode
        if _gpl_ok(gpl_fwd, x, Cin, Cout, math):
            out = _conv_gpl(x, W3, Cout, 1, 0, gpl_fwd, n_out, n_out_dev, Cin, Cout, math)
        else:
            out = torch.empty((n_out, Cout), dtype=torch.float32, device=x.device)
            _lib.call("b200scn_conv_table", x.data_ptr(), x.stride(0), W3.data_ptr(), Cin * Cout, Cout, 1, 0,
                      _lib.ptr(table), table.stride(0) if table is not None else 0, K, n_out, n_out_dev, Cin, Cout,
                      out.data_ptr(), out.stride(0), 0, math, _stream())
        ctx.save_for_backward(x, W3)
        ctx.misc = (table, n_out, n_out_dev, n_in, n_in_dev, symmetric, parent, kofs, W.shape, math, gpl_bwd)
        ctx.gpl_fwd = gpl_fwd
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W3 = ctx.saved_tensors
        table, n_out, n_out_dev, n_in, n_in_dev, symmetric, parent, kofs, w_shape, math, gpl_bwd = ctx.misc
        dout = _rowmajor(dout)
        K, Cin, Cout = W3.shape
        dx = dW = None
        if ctx.needs_input_grad[0] and _gpl_ok(gpl_bwd, dout, Cout, Cin, math):
            # same pair structure, transposed weights; the symmetric (submanifold) list needs the offset flip
            dx = _conv_gpl(dout, W3, 1, Cout, 1 if symmetric else 0, gpl_bwd, n_in, n_in_dev, Cout, Cin, math)
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
            dW = _wgrad(x, dout, table, K, 0, n_out, n_out_dev, w_shape, math, ctx.gpl_fwd, symmetric)
        return dx, dW, None, None, None, None, None, None, None, None, None, None


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
            out = _conv_gpl(x, W3, Cout, 1, 0, gpl_fwd, n_fine, n_fine_dev, Cin, Cout, math)
        else:
            out = torch.empty((n_fine, Cout), dtype=torch.float32, device=x.device)
            _lib.call("b200scn_conv_parent", x.data_ptr(), x.stride(0), W3.data_ptr(), Cin * Cout, Cout, 1,
                      parent.data_ptr(), kofs.data_ptr(), K, n_fine, n_fine_dev, Cin, Cout, out.data_ptr(),
                      out.stride(0), math, _stream())
        ctx.save_for_backward(x, W3)
        ctx.misc = (parent, kofs, child, n_fine, n_fine_dev, n_coarse, n_coarse_dev, math, gpl_bwd)
        ctx.gpl_fwd = gpl_fwd
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W3 = ctx.saved_tensors
        parent, kofs, child, n_fine, n_fine_dev, n_coarse, n_coarse_dev, math, gpl_bwd = ctx.misc
        dout = _rowmajor(dout)
        K, Cin, Cout = W3.shape
        dx = dW = None
        if ctx.needs_input_grad[0] and _gpl_ok(gpl_bwd, dout, Cout, Cin, math):
            dx = _conv_gpl(dout, W3, 1, Cout, 0, gpl_bwd, n_coarse, n_coarse_dev, Cout, Cin, math)
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
                dW = _wgrad(x, dout, child, K, 1, n_fine, n_fine_dev, W3.shape, math, ctx.gpl_fwd, False)
            else:
                dW = _wgrad(x, dout, child, K, 1, n_coarse, n_coarse_dev, W3.shape, math)
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
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        _lib.call("b200scn_bn_bwd_sync", x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), rows, n_dev,
                  rows_hint, C, _lib.ptr(weight), _lib.ptr(bias), mean.data_ptr(), invstd.data_ptr(), leak,
                  int(training), dx.data_ptr(), dx.stride(0), dgamma.data_ptr(), dbeta.data_ptr(), _lib.ptr(sync),
                  scratch.data_ptr(), nbytes, _stream())
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
    reference: uresnet_sparse.py:25,36,44-82).  Returns a tensor [2] = (total_loss, total_acc).  The
    backward pass was already computed by the forward kernel for an upstream gradient of 1 on total_loss;
    `backward` scales it by the incoming gradient of element 0."""

    @staticmethod
    def forward(ctx, feat, W, b, label, ids, nb, pix_w, rows, n_dev, unit_grad):
        _check_cuda_f32(feat, W, b)
        feat = _rowmajor(feat)
        W = W.contiguous()
        label = label.reshape(-1)
        if label.dtype != torch.float32:
            label = label.float()
        if ids.dtype not in (torch.int32, torch.int64, torch.float32, torch.float64):
            ids = ids.long()
        codes = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}
        m, nc = feat.shape[1], W.shape[0]
        L = _lib.load()
        nbytes = L.b200scn_head_loss_scratch_bytes(rows, m, nc, nb)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=feat.device)
        out = torch.empty(2, dtype=torch.float32, device=feat.device)
        need = feat.requires_grad or W.requires_grad
        dfeat = torch.empty((rows, m), dtype=torch.float32, device=feat.device) if need else None
        dW = torch.empty_like(W) if need else None
        db = torch.empty(nc, dtype=torch.float32, device=feat.device) if need else None
        if pix_w is not None:
            pix_w = pix_w.reshape(-1).float()
        _lib.call("b200scn_head_loss", feat.data_ptr(), feat.stride(0), W.data_ptr(), _lib.ptr(b), label.data_ptr(),
                  label.stride(0), _lib.ptr(pix_w), pix_w.stride(0) if pix_w is not None else 0, ids.data_ptr(),
                  codes[ids.dtype], ids.stride(0), rows, n_dev, m, nc, nb, None, 0, out.data_ptr(), _lib.ptr(dfeat),
                  dfeat.stride(0) if dfeat is not None else 0, _lib.ptr(dW), _lib.ptr(db), scratch.data_ptr(), nbytes,
                  _stream())
        ctx.grads = (dfeat, dW, db)
        ctx.unit_grad = unit_grad
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        dfeat, dW, db = ctx.grads
        if not ctx.unit_grad:
            g = gout[0]
            dfeat, dW, db = dfeat * g, dW * g, db * g
        return dfeat, dW, (db if ctx.has_bias else None), None, None, None, None, None, None, None


def head_loss(feat, linear, label, ids, nb, pix_w=None, rows=None, n_dev=None, unit_grad=False):
    """(total_loss, total_acc) of a torch.nn.Linear head + the reference's SegmentationLoss on `feat`."""
    rows = feat.shape[0] if rows is None else rows
    out = HeadLoss.apply(feat, linear.weight, linear.bias, label, ids, nb, pix_w, rows, n_dev, unit_grad)
    return out[0], out[1].detach()
