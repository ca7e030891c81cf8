"""Metadata: per-forward active-site pyramid and rulebook cache.

Plays the role of scn's `Metadata` object that `scn.InputLayer` creates and every later
layer of /root/reference/uresnet/models/uresnet_sparse.py:19-24 consults; all tables are
built on the GPU by libb200scn (csrc/rulebook.cu) and cached per spatial size, so the 8+
submanifold convolutions of a U-Net level share one neighbour table.

Two shape regimes:
  * eager  (default): after each level is built the live row count is read back once, and
    feature matrices have exactly that many rows -- what an scn user expects to see;
  * static (config.static_shapes): nothing is read back; every level keeps the input's row
    count as its capacity and the kernels take the live count from device memory.  This is
    the regime a whole training step is captured in as one CUDA graph.
"""
import weakref

import torch

from .. import _lib
from . import config

_DTYPE_CODE = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Level(object):
    """Active sites of one spatial size."""

    __slots__ = ("cap", "n", "n_dev", "ukeys", "tkeys", "tvals", "slots", "subm", "down", "gpl")

    def __init__(self):
        self.subm = {}   # filter size -> nbr table int32 [K, cap]
        self.down = {}   # stride -> dict(parent, kofs, child, out_size)
        self.gpl = {}    # ("subm", k) / ("child", s) / ("parent", s) -> grouped pair list of the rows of this level

    @property
    def rows(self):
        """Row count feature matrices of this level are allocated with."""
        return self.n if self.n is not None else self.cap


class Metadata(object):
    def __init__(self, dimension):
        self.dimension = dimension
        self.levels = {}
        self.static = config.static_shapes
        self.device = None
        # input-layer maps
        self.n_in = 0
        self.n_in_dev = None  # static regime: device count of valid input rows
        self.in2row = self.first_idx = self.seg_off = self.seg = None
        self._scratch = None
        # rule prefetch (static regime): see _prefetch
        self._input_size = None
        self._plan_key = None
        self._recording = None
        self._prefetching = False
        self._prefetch_event = None

    # -- helpers -----------------------------------------------------------------------
    def _scratch_for(self, n_cap):
        need = _lib.load().b200scn_rulebook_scratch_bytes(n_cap)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._scratch

    def _new_level(self, cap):
        lv = Level()
        lv.cap = cap
        lv.n = None
        lv.slots = _lib.load().b200scn_hash_slots(max(cap, 1))
        dev = self.device
        lv.tkeys = torch.empty(lv.slots, dtype=torch.int64, device=dev)
        lv.tvals = torch.empty(lv.slots, dtype=torch.int32, device=dev)
        lv.ukeys = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
        lv.n_dev = torch.zeros(2, dtype=torch.int32, device=dev)  # [count, error flag]
        return lv

    # -- InputLayer --------------------------------------------------------------------
    def set_input(self, coords, spatial_size):
        """coords [N, D+1] (or [N, D]) any of int32/int64/float32/float64 on a CUDA device."""
        if not coords.is_cuda:
            raise RuntimeError("b200 sparseconvnet: coordinates must live on a CUDA device (no CPU path)")
        if coords.dtype not in _DTYPE_CODE:
            coords = coords.float()
        if coords.dim() != 2 or coords.shape[1] not in (self.dimension, self.dimension + 1):
            raise ValueError("coords must be [N, %d] or [N, %d]" % (self.dimension, self.dimension + 1))
        if coords.shape[0] > 0 and coords.stride(1) != 1:
            coords = coords.contiguous()
        self.device = coords.device
        n = coords.shape[0]
        self.n_in = n
        lv = self._new_level(n)
        dev = self.device
        self.in2row = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.first_idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.seg_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.seg = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        scratch = self._scratch_for(max(n, 1))
        if self.static and config.static_input_rows is not None:
            self.n_in_dev = config.static_input_rows
        _lib.call("b200scn_input_build", _lib.ptr(coords), _DTYPE_CODE[coords.dtype],
                  coords.stride(0) if n > 0 else coords.shape[1],
                  int(coords.shape[1] == self.dimension + 1), self.dimension, n, _lib.ptr(self.n_in_dev),
                  _lib.ptr(lv.tkeys),
                  _lib.ptr(lv.tvals), lv.slots, _lib.ptr(self.in2row), _lib.ptr(lv.ukeys), _lib.ptr(lv.n_dev),
                  _lib.ptr(self.first_idx), _lib.ptr(self.seg_off), _lib.ptr(self.seg),
                  lv.n_dev[1:].data_ptr(), _lib.ptr(scratch), scratch.numel(), _stream())
        if not self.static:
            cnt, err = lv.n_dev.tolist()
            if err:
                raise ValueError("coordinates / sample ids must be integers in [0, 65535)")
            lv.n = cnt
        self.levels[tuple(spatial_size)] = lv
        self._input_size = tuple(spatial_size)
        if self.static and config.prefetch_rules:
            self._plan_key = (self.dimension, self._input_size)
            plan = config.md_plans.get(self._plan_key)
            if plan is None:
                self._recording = config.md_plans[self._plan_key] = []
            elif plan:
                self._prefetch(plan)
        return lv

    # -- rule prefetch --------------------------------------------------------------------
    def _note(self, name, *args):
        """Record a build request of the first static forward (the plan later forwards replay ahead of time)."""
        if self._recording is not None and not self._prefetching:
            ent = (name, args)
            if ent not in self._recording:
                self._recording.append(ent)

    def _prefetch(self, plan):
        """Build everything beyond the input level on the side stream (forked from and later joined to the
        current stream with events, so it is capturable).  Builds that only concern the input level stay lazy on
        the main stream: the first layers need them at once."""
        from . import functional as F
        plan = [(name, args) for name, args in plan
                if not (name in ("subm_table", "subm_gpl") and tuple(args[0]) == self._input_size)]
        if not plan:
            return
        main = torch.cuda.current_stream(self.device)
        side = F._side_stream(self.device)
        side.wait_stream(main)
        self._prefetching = True
        try:
            with torch.cuda.stream(side):
                for name, args in plan:
                    getattr(self, name)(*args)
                self._prefetch_event = side.record_event()
        finally:
            self._prefetching = False

    def _join(self, spatial_size=None):
        """Make the current stream wait for the prefetched builds before it touches anything but the input level."""
        if self._prefetch_event is None or self._prefetching:
            return
        if spatial_size is not None and tuple(spatial_size) == self._input_size:
            return
        torch.cuda.current_stream(self.device).wait_event(self._prefetch_event)
        self._prefetch_event = None

    def check_errors(self):
        """Static regime: the out-of-range flag can be polled after the step (one sync)."""
        for lv in self.levels.values():
            if int(lv.n_dev[1].item()):
                raise ValueError("coordinates / sample ids must be integers in [0, 65535)")

    # -- rulebooks ----------------------------------------------------------------------
    def level(self, spatial_size):
        key = tuple(spatial_size)
        self._join(key)
        if key not in self.levels:
            raise RuntimeError("no active sites registered for spatial size %s" % (key,))
        return self.levels[key]

    def n_dev_ptr(self, lv):
        return lv.n_dev.data_ptr() if lv.n is None else None

    def subm_table(self, spatial_size, filter_size):
        self._note("subm_table", tuple(spatial_size), filter_size)
        lv = self.level(spatial_size)
        if filter_size not in lv.subm:
            K = filter_size ** self.dimension
            rows = lv.rows
            nbr = torch.empty((K, max(rows, 1)), dtype=torch.int32, device=self.device)
            _lib.call("b200scn_subm_table", _lib.ptr(lv.ukeys), rows, self.n_dev_ptr(lv), self.dimension,
                      filter_size, _lib.ptr(lv.tkeys), _lib.ptr(lv.tvals), lv.slots, _lib.ptr(nbr), nbr.stride(0),
                      _stream())
            lv.subm[filter_size] = nbr
        return lv.subm[filter_size]

    def down_rules(self, in_size, out_size, stride):
        """Rules of Convolution(filter == stride): creates the coarser level on first use."""
        self._note("down_rules", tuple(in_size), tuple(out_size), stride)
        self._join()
        lv = self.level(in_size)
        if stride not in lv.down:
            K = stride ** self.dimension
            rows = lv.rows
            nxt = self._new_level(rows)
            parent = torch.empty(max(rows, 1), dtype=torch.int32, device=self.device)
            kofs = torch.empty(max(rows, 1), dtype=torch.uint8, device=self.device)
            child = torch.empty((K, max(rows, 1)), dtype=torch.int32, device=self.device)
            scratch = self._scratch_for(max(rows, 1))
            _lib.call("b200scn_level_down", _lib.ptr(lv.ukeys), rows, self.n_dev_ptr(lv), self.dimension, stride,
                      _lib.ptr(nxt.tkeys), _lib.ptr(nxt.tvals), nxt.slots, _lib.ptr(parent), _lib.ptr(kofs),
                      _lib.ptr(nxt.ukeys), _lib.ptr(nxt.n_dev), _lib.ptr(child), child.stride(0), _lib.ptr(scratch),
                      scratch.numel(), _stream())
            if not self.static:
                nxt.n = int(nxt.n_dev[0].item())
            key = tuple(out_size)
            if key in self.levels:
                raise RuntimeError("spatial size %s already has an active set" % (key,))
            self.levels[key] = nxt
            lv.down[stride] = dict(parent=parent, kofs=kofs, child=child, out_size=key)
        rules = lv.down[stride]
        if rules["out_size"] != tuple(out_size):
            raise RuntimeError("inconsistent coarse spatial size")
        return rules

    # -- grouped pair lists (tensor-core convolution input) ------------------------------------
    def _build_gpl(self, lv, table, parent, kofs, K, size_key=None, layout=0):
        """Pair list over the rows of `lv` (the OUTPUT rows of the convolution that will use it).
        layout: b200scn_gpl_build's (0 = offset-major, read by every kernel; 1 = centre first, large-input
        convolution kernel only -- no per-offset index, so no weight gradient walks it)."""
        L = _lib.load()
        rows = lv.rows
        nd = self.n_dev_ptr(lv)
        tile = (config.gpl_tile_rows or (self.static and config.gpl_tile_rows_by_size.get(size_key))
                or L.b200scn_gpl_tile_rows(max(rows, 1)))
        tiles = max((max(rows, 1) + tile - 1) // tile, 1)
        dev = self.device
        tile_off = torch.empty(tiles + 1, dtype=torch.int32, device=dev)
        # per-(tile, offset) group index
        koff = torch.empty(tiles * (K + 1), dtype=torch.int32, device=dev) if layout == 0 else None
        tab = (_lib.ptr(table), table.stride(0) if table is not None else 0, _lib.ptr(parent), _lib.ptr(kofs), K, rows,
               nd, tile, layout)
        if self.static and rows < config.gpl_fixed_stride_max_rows:
            # static regime: one pass into fixed-stride slots, nothing to scan or read back
            cap = L.b200scn_gpl_max_groups(max(rows, 1), K, tile)
            tile_end = torch.empty(tiles, dtype=torch.int32, device=dev)
            g_rec = torch.empty(cap * 32, dtype=torch.int32, device=dev)
            _lib.call("b200scn_gpl_build", 2, *tab, _lib.ptr(tile_off), _lib.ptr(tile_end), None, _lib.ptr(g_rec), cap,
                      _lib.ptr(koff), None, 0, _stream())
            return dict(tile_off=tile_off, tile_end=tile_end, g_rec=g_rec, K=K, tile_rows=tile, koff=koff, layout=layout,
                        rows_hint=config.gpl_rows_by_size.get(size_key, 0))
        total = torch.zeros(1, dtype=torch.int32, device=dev)
        nbytes = L.b200scn_gpl_scratch_bytes(max(rows, 1))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("b200scn_gpl_build", 0, *tab, _lib.ptr(tile_off), None, _lib.ptr(total), None, 0, None,
                  _lib.ptr(scratch), nbytes, _stream())
        if self.static:
            cap = L.b200scn_gpl_max_groups(max(rows, 1), K, tile)
        else:
            cap = max(int(total.item()), 1)
        g_rec = torch.empty(cap * 32, dtype=torch.int32, device=dev)  # [cap][8] int4 records
        _lib.call("b200scn_gpl_build", 1, *tab, _lib.ptr(tile_off), None, _lib.ptr(total), _lib.ptr(g_rec), cap,
                  _lib.ptr(koff), _lib.ptr(scratch), nbytes, _stream())
        return dict(tile_off=tile_off, tile_end=None, g_rec=g_rec, K=K, tile_rows=tile, koff=koff, layout=layout,
                    rows_hint=config.gpl_rows_by_size.get(size_key, 0) if self.static else 0)

    def subm_gpl(self, spatial_size, filter_size, layout=0):
        """Pair list of the submanifold rules.  The layout-0 list carries `variant(layout)`, which hands out (and
        builds on first use) the centre-first list the large-input convolution kernel prefers."""
        self._note("subm_gpl", tuple(spatial_size), filter_size, layout)
        lv = self.level(spatial_size)
        key = ("subm", filter_size) if layout == 0 else ("subm", filter_size, layout)
        if key not in lv.gpl:
            nbr = self.subm_table(spatial_size, filter_size)
            gpl = self._build_gpl(lv, nbr, None, None, filter_size ** self.dimension, tuple(spatial_size), layout)
            if layout == 0:
                me, size = weakref.ref(self), tuple(spatial_size)
                gpl["variant"] = lambda lay: me().subm_gpl(size, filter_size, lay)
            lv.gpl[key] = gpl
        return lv.gpl[key]

    def strided_gpl(self, in_size, out_size, stride):
        """(pair list over the coarse rows from the child table, pair list over the fine rows from parent/kofs)"""
        self._note("strided_gpl", tuple(in_size), tuple(out_size), stride)
        self._join()
        rules = self.down_rules(in_size, out_size, stride)
        fine, coarse = self.level(in_size), self.level(out_size)
        K = stride ** self.dimension
        if ("child", stride) not in coarse.gpl:
            coarse.gpl[("child", stride)] = self._build_gpl(coarse, rules["child"], None, None, K, tuple(out_size))
        if ("parent", stride) not in fine.gpl:
            fine.gpl[("parent", stride)] = self._build_gpl(fine, None, rules["parent"], rules["kofs"], K, tuple(in_size))
        return coarse.gpl[("child", stride)], fine.gpl[("parent", stride)]

    # -- introspection (tests, debugging) ---------------------------------------------------
    def active_coords(self, spatial_size):
        """int64 [n, D+1] (coords..., sample) of the level's rows, in row order (host sync)."""
        lv = self.level(spatial_size)
        n = lv.n if lv.n is not None else int(lv.n_dev[0].item())
        k = lv.ukeys[:n]
        D = self.dimension
        cols = [(k >> (16 * (D - 1 - d))) & 0xFFFF for d in range(D)]
        cols.append((k >> 48) & 0xFFFF)
        return torch.stack(cols, 1)
