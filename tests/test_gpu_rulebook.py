"""GPU hash-table rulebook vs the oracle: active sets and rule triples bit-exact."""
import numpy as np
import pytest
import torch

from helpers import event_blob, rules_as_set, table_as_set
from oracle import rules

pytestmark = pytest.mark.gpu


def _build(data, dim, ss, levels=3, static=False):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    coords = torch.from_numpy(data[:, : dim + 1]).cuda()
    with scn.static_mode(static):
        md = Metadata(dim)
        size = (ss,) * dim
        md.set_input(coords, size)
        sizes = [size]
        for _ in range(levels - 1):
            nxt = tuple(v // 2 for v in sizes[-1])
            md.down_rules(sizes[-1], nxt, 2)
            sizes.append(nxt)
        for s in sizes:
            md.subm_table(s, 3)
    torch.cuda.synchronize()
    return md, sizes


@pytest.mark.parametrize("dim,dup,dtype", [(3, 0, np.float64), (3, 60, np.float64), (2, 25, np.float32),
                                           (3, 10, np.int64), (3, 10, np.int32)])
def test_rulebook_bit_exact(dim, dup, dtype):
    ss = 32
    data, _ = event_blob(3, ss, dim=dim, dup=dup, seed=5)
    md, sizes = _build(data.astype(dtype), dim, ss)
    c = data[:, : dim + 1].astype(np.int64)
    in2row, act = rules.input_rows(c)
    assert md.levels[sizes[0]].n == len(act)
    assert (md.in2row[: len(c)].cpu().numpy() == in2row).all()
    for lvl, size in enumerate(sizes):
        lv = md.levels[size]
        got = md.active_coords(size).cpu().numpy()
        assert got.shape == act.shape and (got == act).all(), "level %d active set / order" % lvl
        nbr = md.subm_table(size, 3)
        assert table_as_set(nbr, lv.n) == rules_as_set(rules.subm_rules(act)), "level %d subm rules" % lvl
        if lvl + 1 < len(sizes):
            parent, kofs, coarse = rules.down_rules(act)
            r = lv.down[2]
            assert (r["parent"][: lv.n].cpu().numpy() == parent).all()
            assert (r["kofs"][: lv.n].cpu().numpy() == kofs).all()
            child = r["child"][:, : len(coarse)].cpu().numpy()
            exp = np.full_like(child, -1)
            exp[kofs, parent] = np.arange(len(parent))
            assert (child == exp).all()
            act = coarse


def test_duplicate_csr_is_in_input_order():
    data, _ = event_blob(1, 16, dup=200, seed=9)
    md, sizes = _build(data, 3, 16, levels=1)
    n = md.levels[sizes[0]].n
    in2row = md.in2row[: len(data)].cpu().numpy()
    first = md.first_idx[:n].cpu().numpy()
    off = md.seg_off[: n + 1].cpu().numpy()
    seg = md.seg[: off[-1]].cpu().numpy()
    assert off[-1] == len(data) - n
    for r in range(n):
        members = np.nonzero(in2row == r)[0]
        assert first[r] == members[0]
        assert (seg[off[r]:off[r + 1]] == members[1:]).all()


def test_static_regime_matches_eager():
    data, _ = event_blob(2, 32, dup=30, seed=6)
    md_e, sizes = _build(data, 3, 32, static=False)
    md_s, _ = _build(data, 3, 32, static=True)
    for size in sizes:
        le, ls = md_e.levels[size], md_s.levels[size]
        assert ls.n is None and int(ls.n_dev[0].item()) == le.n
        assert (md_s.active_coords(size) == md_e.active_coords(size)).all()
        assert table_as_set(ls.subm[3], le.n) == table_as_set(le.subm[3], le.n)


def test_empty_and_single_point_inputs():
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    md = Metadata(3)
    lv = md.set_input(torch.zeros((0, 4), dtype=torch.float64, device="cuda"), (16, 16, 16))
    assert lv.n == 0
    md = Metadata(3)
    lv = md.set_input(torch.tensor([[3, 4, 5, 0]], dtype=torch.int64, device="cuda"), (16, 16, 16))
    assert lv.n == 1
    nbr = md.subm_table((16, 16, 16), 3)
    t = nbr[:, :1].cpu().numpy().reshape(-1)
    assert t[13] == 0 and (np.delete(t, 13) == -1).all()


def test_out_of_range_coordinates_raise():
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    md = Metadata(3)
    with pytest.raises(ValueError):
        md.set_input(torch.tensor([[3, -4, 5, 0]], dtype=torch.int64, device="cuda"), (16, 16, 16))


def test_large_event_properties():
    """BASELINE size (512^3, several events): size-independent invariants instead of the oracle."""
    data, _ = event_blob(4, 512, seed=1)
    md, sizes = _build(data, 3, 512, levels=5)
    prev = None
    for size in sizes:
        lv = md.levels[size]
        c = md.active_coords(size)
        assert len(torch.unique(c, dim=0)) == lv.n  # no duplicate sites
        nbr = lv.subm[3][:, : lv.n]
        K = nbr.shape[0]
        assert (nbr[K // 2] == torch.arange(lv.n, device="cuda")).all()
        # symmetry of the rulebook: nbr[K-1-k][nbr[k][o]] == o
        for k in range(K // 2):
            o = torch.nonzero(nbr[k] >= 0)[:, 0]
            assert (nbr[K - 1 - k][nbr[k][o].long()] == o).all()
        if prev is not None:
            pc, rules_ = prev
            par = rules_["parent"][: len(pc)].long()
            assert (c[par][:, :3] == pc[:, :3] // 2).all() and (c[par][:, 3] == pc[:, 3]).all()
        if 2 in lv.down:
            prev = (c, lv.down[2])
