"""Tensor-core (mma.sync TF32) kernels vs the CPU oracle.

MATH_TF32X3 (3-term operand split) must be fp32-equivalent: same tolerances as the FFMA
kernels.  MATH_TF32 (single pass, 10-bit mantissa operands) is the reduced-precision mode:
max|d| <= 5e-3 * scale on outputs, rel-L2 <= 1e-2 on gradients."""
import os

import pytest
import torch

from conftest import small_flags
from helpers import check_grads, event_blob, rel_l2
from oracle import scn_cpu as ocn
from oracle.model import OracleUResNet, segmentation_loss
from test_gpu_layers import _check, _lift, _pair

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[(True, 0), (False, 0), (False, 1 << 30)], ids=["tc", "gpl", "pairs"])
def _conv_path(request):
    """Run every test of this file through the three tensor-core conv paths: tcgen05 rule-table implicit GEMM
    (sparse_tc.cu; the default), mma.sync over grouped pair lists for all levels (threshold 0) and mma.sync with
    in-kernel listing only (threshold huge)."""
    from uresnet_pytorch_b200.sparseconvnet import config
    prev = config.gpl_min_rows, config.use_tc, config.tc_min_channels, config.tc_wgrad_min_channels
    config.use_tc, config.gpl_min_rows = request.param
    config.tc_min_channels = config.tc_wgrad_min_channels = 16   # every width through the tcgen05 kernels in "tc"
    yield
    config.gpl_min_rows, config.use_tc, config.tc_min_channels, config.tc_wgrad_min_channels = prev

TOL = {1: (3e-5, 2e-4), 2: (5e-3, 1e-2)}  # math mode -> (forward, gradient)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("dim,cin,cout", [(3, 16, 16), (3, 32, 16), (3, 32, 32), (3, 48, 48), (3, 64, 64),
                                          (3, 16, 48), (2, 16, 32), (3, 96, 48)])
def test_submanifold_convolution_mma(mode, dim, cin, cout):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(dim, 24, dup=15)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, cin)
    torch.manual_seed(2)
    with scn.math(mode):
        _check(ocn.SubmanifoldConvolution(dim, cin, cout, 3, False),
               scn.SubmanifoldConvolution(dim, cin, cout, 3, False), x_o, x_g, f_o, f_g, *TOL[mode])


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("dim,cin,cout", [(3, 16, 32), (3, 64, 80), (2, 32, 48)])
def test_strided_mma(mode, dim, cin, cout):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(dim, 24, dup=15)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, cin)
    torch.manual_seed(2)
    seq_o = ocn.Sequential().add(ocn.Convolution(dim, cin, cout, 2, 2, False)).add(
        ocn.Deconvolution(dim, cout, cin, 2, 2, False))
    seq_g = scn.Sequential().add(scn.Convolution(dim, cin, cout, 2, 2, False)).add(
        scn.Deconvolution(dim, cout, cin, 2, 2, False))
    with scn.math(mode):
        _check(seq_o, seq_g, x_o, x_g, f_o, f_g, *TOL[mode])


@pytest.mark.parametrize("mode", [1, 2])
def test_network_in_network_mma(mode):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(3, 24)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, 32)
    torch.manual_seed(2)
    with scn.math(mode):
        _check(ocn.NetworkInNetwork(32, 16, False), scn.NetworkInNetwork(32, 16, False), x_o, x_g, f_o, f_g,
               *TOL[mode])


@pytest.mark.parametrize("mode,logit_tol,grad_tol", [(1, 1e-4, 1e-3), (2, 2e-2, 3e-2)])
def test_model_mma_matches_oracle(mode, logit_tol, grad_tol):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from test_gpu_model import _loss_gpu, _models
    flags = small_flags(SPATIAL_SIZE=64, URESNET_FILTERS=16, URESNET_NUM_STRIDES=4)
    ref, net = _models(flags)
    data, label = event_blob(2, 64, dup=25)
    d, l = torch.from_numpy(data), torch.from_numpy(label)
    out_ref = ref(d)[0]
    loss_ref, _ = segmentation_loss(out_ref, d, l)
    loss_ref.backward()
    with scn.math(mode):
        out, loss, _ = _loss_gpu(net, d.cuda(), l.cuda(), 2)
        loss.backward()
    torch.cuda.synchronize()
    scale = out_ref.abs().max().item()
    assert (out.detach().cpu() - out_ref.detach()).abs().max().item() <= logit_tol * scale
    check_grads(net, ref, grad_tol, 20 * grad_tol)


def test_mma_large_batch_against_fp32_kernel():
    """BASELINE-size check: tf32x3 tensor-core conv == FFMA conv on a 512^3 multi-event batch."""
    from uresnet_pytorch_b200 import sparseconvnet as scn
    data, _ = event_blob(8, 512, seed=2)
    coords = torch.from_numpy(data[:, :4]).cuda()
    torch.manual_seed(0)
    feats = torch.randn(len(data), 1, device="cuda")
    x = scn.InputLayer(3, 512, mode=3)((coords, feats))
    lift = scn.NetworkInNetwork(1, 16, False).cuda()
    conv = scn.SubmanifoldConvolution(3, 16, 16, 3, False).cuda()
    with torch.no_grad():
        h = lift(x)
        with scn.math(0):
            a = conv(h).features
        with scn.math(1):
            b = conv(h).features
        with scn.math(2):
            c = conv(h).features
    assert (a - b).abs().max().item() <= 2e-5 * a.abs().max().item()
    assert (a - c).abs().max().item() <= 5e-3 * a.abs().max().item()


@pytest.mark.parametrize("tile_rows", [32, 64, 128])
def test_grouped_pair_lists_match_table(tile_rows, monkeypatch):
    """The GPL builder reproduces exactly the (offset, in, out) triples of the neighbour table."""
    import numpy as np
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    from helpers import table_as_set
    data, _ = event_blob(3, 48, dup=20, seed=12)
    coords = torch.from_numpy(data[:, :4]).cuda()
    monkeypatch.setattr(scn.config, "gpl_tile_rows", tile_rows)
    md = Metadata(3)
    size = (48, 48, 48)
    lv = md.set_input(coords, size)
    gpl = md.subm_gpl(size, 3)
    off = gpl["tile_off"].cpu().numpy()
    R = gpl["tile_rows"]
    rec = gpl["g_rec"].cpu().numpy().reshape(-1, 8, 4)   # [group][entry j][in(j), in(j+8), out bytes | k, 0]
    triples = set()
    for tile in range(len(off) - 1):
        ks = []
        for grp in range(off[tile], off[tile + 1]):
            kset = set()
            for j in range(8):
                word = int(rec[grp, j, 2]) & 0xFFFFFFFF
                kset.add((word >> 16) & 0xFF)
                assert rec[grp, j, 3] == 0 and (word >> 24) == 0
                for half in range(2):
                    o = (word >> (8 * half)) & 0xFF
                    if o < R:
                        triples.add(((word >> 16) & 0xFF, int(rec[grp, j, half]), tile * R + o))
                    else:
                        assert o == R and rec[grp, j, half] == 0   # padding: dummy output row
            assert len(kset) == 1                   # one offset per group
            ks.append(kset.pop())
        assert ks == sorted(ks)                     # groups of a tile are ordered by offset
    assert triples == table_as_set(md.subm_table(size, 3), lv.n)
    assert off[-1] == len(rec)


def test_centre_first_pair_list_matches_table(monkeypatch):
    """Layout 1 (what the large-input kernel reads): the same (offset, in, out) triples as the neighbour table; every
    tile starts with its 8 centre groups -- slot j of group s is output row 16 s + j, flag "C = 0" -- followed by the
    other offsets in ascending order without flags."""
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    from helpers import table_as_set
    data, _ = event_blob(3, 48, dup=20, seed=12)
    coords = torch.from_numpy(data[:, :4]).cuda()
    monkeypatch.setattr(scn.config, "gpl_tile_rows", 128)
    md = Metadata(3)
    size = (48, 48, 48)
    lv = md.set_input(coords, size)
    gpl = md.subm_gpl(size, 3, 1)
    assert gpl["layout"] == 1 and gpl["koff"] is None and gpl["tile_rows"] == 128
    off = gpl["tile_off"].cpu().numpy()
    rec = gpl["g_rec"].cpu().numpy().reshape(-1, 8, 4)
    n, R, centre = lv.n, 128, 13
    triples = set()
    for tile in range(len(off) - 1):
        assert off[tile + 1] - off[tile] >= R // 16
        ks = []
        for gi, grp in enumerate(range(off[tile], off[tile + 1])):
            dense = gi < R // 16
            for j in range(8):
                word = int(rec[grp, j, 2]) & 0xFFFFFFFF
                k, flags = (word >> 16) & 0xFF, word >> 24
                assert rec[grp, j, 3] == 0 and flags == (1 if dense else 0)
                assert (k == centre) == dense
                for half in range(2):
                    o = (word >> (8 * half)) & 0xFF
                    if dense:
                        assert o == 16 * gi + j + 8 * half                 # the slot IS the output row
                        row = tile * R + o
                        if row < n:
                            assert rec[grp, j, half] == row                # centre rule: in == out
                            triples.add((k, row, row))
                        else:
                            assert rec[grp, j, half] == 0                  # past the end: any valid input row
                    elif o < R:
                        triples.add((k, int(rec[grp, j, half]), tile * R + o))
                    else:
                        assert o == R and rec[grp, j, half] == 0
            if not dense and j == 7:
                ks.append(k)
        assert ks == sorted(ks)
    assert triples == table_as_set(md.subm_table(size, 3), n)
    assert off[-1] == len(rec)


@pytest.mark.parametrize("mode", [1, 2])
def test_gpl_and_pairs_kernels_agree(mode):
    """Same layer through the grouped-pair-list kernel and the in-kernel-listing kernel."""
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(3, 24, dup=15)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, 32)
    torch.manual_seed(3)
    conv = scn.SubmanifoldConvolution(3, 32, 48, 3, False).cuda()
    with torch.no_grad(), scn.math(mode):
        a = conv(x_g).features
        prev = scn.config.use_gpl, scn.config.use_tc
        scn.config.use_gpl = scn.config.use_tc = False
        try:
            b = conv(x_g).features
        finally:
            scn.config.use_gpl, scn.config.use_tc = prev
    assert (a - b).abs().max().item() <= (1e-5 if mode == 1 else 2e-3) * a.abs().max().item()


@pytest.mark.parametrize("mode,cin,cout,layout", [(1, 16, 16, 0), (1, 16, 16, 1), (2, 16, 16, 1), (1, 64, 32, 1),
                                                  (1, 32, 32, 1), (1, 16, 32, 1), (2, 32, 64, 1)])
def test_large_input_gpl_kernel(mode, cin, cout, layout):
    """k_conv_gpl (one warp per 128-row tile, weights resident in shared memory -- the kernel bench.py's
    roofline times) against the neighbour-table kernel, for both pair-list layouts (offset-major, centre first);
    then the module path (SubmanifoldConvolution forward, dgrad and
    wgrad with the list variant functional._conv_gpl picks) against the FFMA mode.  Runs in a subprocess because
    the row threshold that selects the kernel is read once per process (B200SCN_GPL_LARGE_MIN_ROWS)."""
    import subprocess
    import sys
    from uresnet_pytorch_b200.sparseconvnet import config
    if config.use_tc or config.gpl_min_rows != 0:
        pytest.skip("subprocess test: one run (under the 'gpl' path id) is enough")
    code = r"""
import torch, sys
sys.path.insert(0, %r)
from helpers import event_blob
from uresnet_pytorch_b200 import sparseconvnet as scn, _lib
from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
scn.config.gpl_tile_rows = 128
data, _ = event_blob(6, 96, seed=21)
coords = torch.from_numpy(data[:, :4]).cuda()
md = Metadata(3); size = (96, 96, 96)
lv = md.set_input(coords, size)
nbr = md.subm_table(size, 3); layout = %d
n, K, Cin, Cout, mode = lv.n, 27, %d, %d, %d
torch.manual_seed(1)
x = torch.randn(n, Cin, device="cuda"); W = torch.randn(K, Cin, Cout, device="cuda") * 0.1
st = torch.cuda.current_stream().cuda_stream
served = 0
for (si, so, flip, ci, co) in ((Cout, 1, 0, Cin, Cout), (1, Cout, 1, Cout, Cin)):
    # what the library serves for this direction: 0 = not the large-input kernel (shared-memory budget)
    want = _lib.load().b200scn_conv_gpl_layout(128, 1, n, ci, co, K, 1, mode)
    lay = 0 if want == 0 else layout
    served += want != 0
    gpl = md.subm_gpl(size, 3, lay)
    assert gpl["tile_rows"] == 128 and gpl["tile_end"] is None and gpl["layout"] == lay
    xin = x if ci == Cin else torch.randn(n, ci, device="cuda")
    a = torch.empty(n, co, device="cuda"); b = torch.empty(n, co, device="cuda")
    _lib.call("b200scn_conv_gpl", xin.data_ptr(), ci, W.data_ptr(), None, Cin * Cout, si, so, flip, 128, gpl["layout"],
              gpl["tile_off"].data_ptr(), None, gpl["g_rec"].data_ptr(), K, n, None, 0, ci, co, a.data_ptr(), co, 0, None, 0,
              mode, st)
    _lib.call("b200scn_conv_table", xin.data_ptr(), ci, W.data_ptr(), Cin * Cout, si, so, flip, nbr.data_ptr(),
              nbr.stride(0), K, n, None, ci, co, b.data_ptr(), co, 0, 0, st)
    torch.cuda.synchronize()
    tol = 3e-5 if mode == 1 else 5e-3
    err = (a - b).abs().max().item() / b.abs().max().item()
    assert err <= tol, err
assert served or Cin > 32, "the large-input kernel must serve the narrow shapes"
# module path: forward, dgrad, wgrad with the variant list the library asks for
scn.config.use_tc = False
coords_l = torch.from_numpy(data[:, :4]).long().cuda()
feats = torch.randn(len(data), Cin, device="cuda")
conv = scn.SubmanifoldConvolution(3, Cin, Cout, 3, False).cuda()
res = {}
for m in (0, mode):
    f = feats.clone().requires_grad_(True)
    conv.weight.grad = None
    with scn.math(m):
        xs = scn.InputLayer(3, torch.LongTensor(size), 3)((coords_l, f))
        y = conv(xs)
        assert m == 0 or not served or any(len(k) == 3 for k in y.metadata.levels[size].gpl), y.metadata.levels[size].gpl.keys()
        (y.features * torch.linspace(-1, 1, Cout, device="cuda")).sum().backward()
    res[m] = (y.features.detach(), f.grad.detach(), conv.weight.grad.detach().clone())
for a, b in zip(res[0], res[mode]):
    err = (a - b).abs().max().item() / a.abs().max().item()
    assert err <= (1e-4 if mode == 1 else 2e-2), err
print("ok", n)
""" % (os.path.dirname(os.path.abspath(__file__)), layout, cin, cout, mode)
    env = dict(os.environ, B200SCN_GPL_LARGE_MIN_ROWS="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("cin,cout", [(16, 16), (48, 80), (80, 48), (160, 96)])
def test_conv_tc_residual_and_device_row_count(cin, cout):
    """b200scn_conv_tc straight through the C-ABI: live row count from device memory below the buffer capacity
    (rows beyond it neither read nor written), fused residual add, channel counts that are not multiples of 32,
    output passes (Cout > 128 is covered by the c5 trainer test), against the FFMA neighbour-table kernel."""
    from uresnet_pytorch_b200 import _lib
    from uresnet_pytorch_b200.sparseconvnet import functional as F
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    data, _ = event_blob(3, 64, seed=17)
    coords = torch.from_numpy(data[:, :4]).cuda()
    md = Metadata(3)
    size = (64, 64, 64)
    lv = md.set_input(coords, size)
    nbr = md.subm_table(size, 3)
    n, K = lv.n, 27
    cap = n + 300
    torch.manual_seed(1)
    x = torch.randn(cap, cin, device="cuda")
    W = torch.nn.Parameter(torch.randn(K, cin, cout, device="cuda") * 0.1)
    res = torch.randn(cap, cout, device="cuda")
    table = torch.full((K, cap), -1, dtype=torch.int32, device="cuda")
    table[:, :n] = nbr[:, :n]
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ref = torch.empty(n, cout, device="cuda")
    _lib.call("b200scn_conv_table", x.data_ptr(), cin, W.data_ptr(), cin * cout, cout, 1, 0, nbr.data_ptr(),
              nbr.stride(0), K, n, None, cin, cout, ref.data_ptr(), cout, 0, 0, st)
    ent = F.packed_weights(W, W.detach(), cin, cout, cout, 1, 0, "tc")
    hi = ent["buf"].data_ptr()
    # 3xTF32 vs FFMA differ by summation order only: ~sqrt(27 cin) ulps of the largest partial sum
    for mode, tol in ((1, 3e-5 if cin <= 80 else 6e-5), (2, 5e-3)):
        out = torch.full((cap, cout), 7.0, device="cuda")
        _lib.call("b200scn_conv_tc", x.data_ptr(), cin, hi, hi + 4 * ent["elems"], table.data_ptr(), cap, K, cap,
                  n_dev.data_ptr(), 0, cin, cout, out.data_ptr(), cout, res.data_ptr(), cout, mode, st)
        torch.cuda.synchronize()
        want = ref + res[:n]
        assert (out[:n] - want).abs().max().item() <= tol * want.abs().max().item()
        assert (out[n:] == 7.0).all()


@pytest.mark.parametrize("cin,cout,K", [(16, 16, 27), (32, 48, 27), (48, 80, 27), (80, 64, 27), (64, 32, 8), (32, 16, 1),
                                        (96, 160, 27)])
def test_wgrad_tc_matches_ffma_kernel(cin, cout, K):
    """b200scn_wgrad_tc (tcgen05, MN-major operands, row splits folded in fixed order) through the C-ABI against
    the FFMA table kernel: device-side row count below the capacity with garbage (NaN) beyond it, accumulation
    into an existing buffer, submanifold (27), strided (8, child table) and NetworkInNetwork (1, no table) rules."""
    from uresnet_pytorch_b200 import _lib
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    data, _ = event_blob(3, 64, seed=19)
    coords = torch.from_numpy(data[:, :4]).cuda()
    md = Metadata(3)
    size = (64, 64, 64)
    lv = md.set_input(coords, size)
    L = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    if K == 27:
        tab, n_out, n_in = md.subm_table(size, 3), lv.n, lv.n
    elif K == 8:
        rules = md.down_rules(size, (32, 32, 32), 2)
        tab, n_out, n_in = rules["child"], md.level((32, 32, 32)).n, lv.n
    else:
        tab, n_out, n_in = None, lv.n, lv.n
    cap = n_out + 200
    torch.manual_seed(4)
    x = torch.randn(n_in, cin, device="cuda")
    dy = torch.full((cap, cout), float("nan"), device="cuda")
    dy[:n_out] = torch.randn(n_out, cout, device="cuda")
    table = None
    if tab is not None:
        table = torch.full((K, cap), -1, dtype=torch.int32, device="cuda")
        table[:, :n_out] = tab[:, :n_out]
    n_dev = torch.tensor([n_out], dtype=torch.int32, device="cuda")
    ref = torch.empty(K, cin, cout, device="cuda")
    nb = L.b200scn_wgrad_scratch_bytes(n_out, K, cin, cout)
    scr = torch.empty(nb, dtype=torch.uint8, device="cuda")
    _lib.call("b200scn_wgrad_table", x.data_ptr(), cin, dy.data_ptr(), cout, _lib.ptr(tab), tab.stride(0) if tab is not None else 0,
              K, 0, n_out, None, cin, cout, ref.data_ptr(), 0, 0, scr.data_ptr(), nb, st)
    for mode, tol in ((1, 5e-5), (2, 5e-3)):
        for hint in (0, n_out):
            nb2 = L.b200scn_wgrad_tc_scratch_bytes(hint if hint else cap, K, cin, cout)
            scr2 = torch.empty(nb2, dtype=torch.uint8, device="cuda")
            base = torch.randn(K, cin, cout, device="cuda")
            dW = base.clone()
            _lib.call("b200scn_wgrad_tc", x.data_ptr(), cin, dy.data_ptr(), cout, _lib.ptr(table), cap if table is not None else 0,
                      K, cap, n_dev.data_ptr(), hint, cin, cout, dW.data_ptr(), 1, mode, scr2.data_ptr(), nb2, st)
            torch.cuda.synchronize()
            err = (dW - base - ref).abs().max().item() / ref.abs().max().item()
            assert err <= tol, (mode, hint, err)
