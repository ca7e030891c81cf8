"""ctypes binding of libb200scn.so (the C-ABI declared in include/b200scn.h).

There is no CPU fallback: if the library is missing the first call raises.
"""
import ctypes
import os

from ._build import SO_PATH

_c = ctypes
_I64, _I32, _F32, _P = _c.c_int64, _c.c_int, _c.c_float, _c.c_void_p

# name -> (restype, argtypes); kept in lock-step with include/b200scn.h (tests check the
# exported symbol list against the header).
SIGNATURES = {
    "b200scn_last_error": (_c.c_char_p, []),
    "b200scn_version": (_I32, []),
    "b200scn_sm_count": (_I32, []),
    "b200scn_launch_count": (_I64, []),
    "b200scn_hash_slots": (_I64, [_I64]),
    "b200scn_rulebook_scratch_bytes": (_I64, [_I64]),
    "b200scn_input_build": (_I32, [_P, _I32, _I64, _I32, _I32, _I64, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P,
                                   _P, _I64, _P]),
    "b200scn_level_down": (_I32, [_P, _I64, _P, _I32, _I32, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P, _I64, _P]),
    "b200scn_subm_table": (_I32, [_P, _I64, _P, _I32, _I32, _P, _P, _I64, _P, _I64, _P]),
    "b200scn_gpl_tile_rows": (_I32, [_I64]),
    "b200scn_gpl_max_groups": (_I64, [_I64, _I32, _I32]),
    "b200scn_gpl_scratch_bytes": (_I64, [_I64]),
    "b200scn_gpl_build": (_I32, [_I32, _P, _I64, _P, _P, _I32, _I64, _P, _I32, _I32, _P, _P, _P, _P, _I64, _P, _P, _I64,
                                 _P]),
    "b200scn_conv_gpl_layout": (_I32, [_I32, _I32, _I64, _I32, _I32, _I32, _I32, _I32]),
    "b200scn_wgrad_gpl_scratch_bytes": (_I64, [_I32, _I32, _I32]),
    "b200scn_wgrad_gpl": (_I32, [_P, _I64, _P, _I64, _I32, _P, _P, _I32, _I32, _I64, _P, _I64, _I32, _I32, _P, _I32,
                                 _I32, _P, _I64, _P]),
    "b200scn_conv_gpl": (_I32, [_P, _I64, _P, _P, _I64, _I64, _I64, _I32, _I32, _I32, _P, _P, _P, _I32, _I64, _P, _I64,
                                _I32, _I32, _P, _I64, _I32, _P, _I64, _I32, _P]),
    "b200scn_pack_weights": (_I32, [_P, _I32, _I64, _P]),
    "b200scn_conv_tc_packed_elems": (_I64, [_I32, _I32, _I32]),
    "b200scn_conv_tc_supported": (_I32, [_I32, _I32, _I32]),
    "b200scn_pack_weights_tc": (_I32, [_P, _I32, _I64, _P]),
    "b200scn_conv_tc": (_I32, [_P, _I64, _P, _P, _P, _I64, _I32, _I64, _P, _I64, _I32, _I32, _P, _I64, _P, _I64, _I32,
                               _P]),
    "b200scn_wgrad_tc_scratch_bytes": (_I64, [_I64, _I32, _I32, _I32]),
    "b200scn_wgrad_tc": (_I32, [_P, _I64, _P, _I64, _P, _I64, _I32, _I64, _P, _I64, _I32, _I32, _P, _I32, _I32, _P,
                                _I64, _P]),
    "b200scn_dense_conv_tc_scratch_bytes": (_I64, [_I32, _I32, _I32]),
    "b200scn_dense_conv_tc": (_I32, [_P, _I64, _P, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _P, _I64, _I32, _P, _I64,
                                     _P]),
    "b200scn_adam_step": (_I32, [_P, _I32, _I64, _P, _P, _F32, _F32, _F32, _F32, _P]),
    "b200scn_head_loss_scratch_bytes": (_I64, [_I64, _I32, _I32, _I32]),
    "b200scn_head_loss": (_I32, [_P, _I64, _P, _P, _P, _I64, _P, _I64, _P, _I32, _I64, _I64, _P, _I32, _I32, _I32, _P,
                                 _I64, _P, _P, _I64, _P, _P, _I32, _P, _I64, _P]),
    "b200scn_head_predict_scratch_bytes": (_I64, [_I64, _I32, _I32]),
    "b200scn_head_predict": (_I32, [_P, _I64, _P, _P, _P, _I64, _P, _I32, _I64, _P, _I32, _I64, _I64, _P, _I32, _I32,
                                    _I32, _P, _I64, _P, _I64, _P, _P, _P, _I64, _P]),
    "b200scn_conv_table": (_I32, [_P, _I64, _P, _I64, _I64, _I64, _I32, _P, _I64, _I32, _I64, _P, _I32, _I32, _P,
                                  _I64, _I32, _I32, _P]),
    "b200scn_conv_parent": (_I32, [_P, _I64, _P, _I64, _I64, _I64, _P, _P, _I32, _I64, _P, _I32, _I32, _P, _I64,
                                   _I32, _P]),
    "b200scn_wgrad_scratch_bytes": (_I64, [_I64, _I32, _I32, _I32]),
    "b200scn_wgrad_table": (_I32, [_P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I64, _P, _I32, _I32, _P, _I32, _I32,
                                   _P, _I64, _P]),
    "b200scn_bn_scratch_bytes": (_I64, [_I64, _I32]),
    "b200scn_bn_fwd_train": (_I32, [_P, _I64, _I64, _P, _I32, _P, _P, _P, _P, _F32, _F32, _F32, _P, _I64, _P, _P,
                                    _P, _I64, _P]),
    "b200scn_bn_fwd_eval": (_I32, [_P, _I64, _I64, _P, _I32, _P, _P, _P, _P, _F32, _F32, _P, _I64, _P]),
    "b200scn_bn_bwd": (_I32, [_P, _I64, _P, _I64, _I64, _P, _I32, _P, _P, _P, _P, _F32, _I32, _P, _I64, _P, _P, _P,
                              _I64, _P]),
    "b200scn_bn_fwd_train_sync": (_I32, [_P, _I64, _I64, _P, _I64, _I32, _P, _P, _P, _P, _F32, _F32, _F32, _P, _I64, _P,
                                         _P, _P, _P, _I64, _P]),
    "b200scn_bn_bwd_sync": (_I32, [_P, _I64, _P, _I64, _I64, _P, _I64, _I32, _P, _P, _P, _P, _F32, _I32, _P, _I64, _P,
                                   _P, _P, _P, _I64, _P]),
    "b200scn_collate": (_I32, [_P, _I64, _P, _I32, _I32, _I32, _I64, _P, _I64, _P, _P]),
    "b200scn_gather_rows": (_I32, [_P, _I64, _P, _I64, _P, _I32, _P, _I64, _P]),
    "b200scn_segsum_rows": (_I32, [_P, _I64, _P, _P, _P, _I64, _P, _I32, _P, _I64, _P]),
    "b200scn_dense_conv": (_I32, [_P, _I64, _P, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _P, _I64, _P]),
    "b200scn_dense_wgrad_scratch_bytes": (_I64, [_P, _I32, _I32, _I32]),
    "b200scn_dense_wgrad": (_I32, [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _I32, _P, _I32, _P, _I64, _P]),
    "b200scn_pad_replicate": (_I32, [_P, _P, _I32, _I32, _P, _P, _P, _I32, _P]),
    "b200scn_colsum_scratch_bytes": (_I64, [_I32]),
    "b200scn_colsum": (_I32, [_P, _I64, _I64, _I32, _P, _P, _I64, _P]),
}

_lib = None


def load():
    """dlopen the in-tree library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "libb200scn.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "this package has no CPU or PyTorch fallback." % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().b200scn_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Invoke an int-returning entry point; raise RuntimeError(last_error) on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, last_error()))


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else t.data_ptr()
