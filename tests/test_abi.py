"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/b200scn.h declares, the ctypes table agrees with the header, and the Python module
surface has the scn names / state-dict layout the reference model needs.  No compute calls."""
import ctypes
import importlib.util
import os
import re
import sys

import pytest
import torch

from conftest import ROOT, small_flags

HEADER = os.path.join(ROOT, "include", "b200scn.h")


def _header_decls():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(b200scn_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        decls[m.group(1)] = n
    return decls


def test_library_exports_every_declared_symbol():
    from uresnet_pytorch_b200 import _build, _lib
    assert os.path.exists(_build.SO_PATH)
    lib = ctypes.CDLL(_build.SO_PATH)
    decls = _header_decls()
    assert len(decls) >= 19
    for name in decls:
        assert hasattr(lib, name), "library does not export %s" % name
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, nargs in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name
    L = _lib.load()
    assert L.b200scn_version() >= 1
    assert L.b200scn_hash_slots(1000) == 2048
    assert L.b200scn_rulebook_scratch_bytes(1000) > 0
    assert L.b200scn_wgrad_scratch_bytes(1000, 27, 16, 16) >= 27 * 16 * 16 * 4
    assert L.b200scn_bn_scratch_bytes(1000, 16) > 0


def test_argument_errors_are_reported_not_thrown():
    from uresnet_pytorch_b200 import _lib
    L = _lib.load()
    # table == NULL with K != 1 is an argument error caught before any CUDA call
    rc = L.b200scn_conv_table(None, 16, None, 256, 16, 1, 0, None, 0, 27, 0, None, 16, 16, None, 16, 0, 0, None)
    assert rc == -1 and "K == 1" in _lib.last_error()
    with pytest.raises(RuntimeError):
        _lib.call("b200scn_subm_table", None, 10, None, 3, 4, None, None, 32, None, 10, None)


def test_sm100a_cubin_only():
    """The library carries sm_100a SASS (no PTX JIT, no other arch)."""
    import subprocess
    from uresnet_pytorch_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", _build.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_module_surface_and_state_dict_layout():
    from oracle.model import OracleUResNet
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.model import UResNet
    for name in ["Sequential", "InputLayer", "SubmanifoldConvolution", "Convolution", "Deconvolution", "UNet",
                 "BatchNormReLU", "BatchNormLeakyReLU", "NetworkInNetwork", "ConcatTable", "AddTable", "JoinTable",
                 "Identity", "OutputLayer", "SparseConvNetTensor", "Metadata"]:
        assert hasattr(scn, name), name
    flags = small_flags(URESNET_FILTERS=16, URESNET_NUM_STRIDES=5)
    torch.manual_seed(0)
    net, ref = UResNet(flags), OracleUResNet(flags)
    a, b = net.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    assert sum(p.numel() for p in net.parameters()) == 2741477
    convs = [m for m in net.modules() if isinstance(m, scn.SubmanifoldConvolution)]
    assert len(convs) == 37 and convs[0].weight.shape == (27, 1, 16)
    assert len([m for m in net.modules() if isinstance(m, scn.Convolution)]) == 4
    assert len([m for m in net.modules() if isinstance(m, scn.Deconvolution)]) == 4
    assert len([m for m in net.modules() if isinstance(m, scn.NetworkInNetwork)]) == 4
    assert len([m for m in net.modules() if isinstance(m, scn.BatchNormalization)]) == 45
    # newer scn checkpoints carry a groups dimension on conv weights
    sd = {k: (v.unsqueeze(1) if v.dim() == 3 else v) for k, v in b.items()}
    net.load_state_dict(sd)
    assert torch.equal(net.sparseModel[1].weight, ref.sparseModel[1].weight)


def test_cpu_tensors_fail_loudly():
    from uresnet_pytorch_b200.model import UResNet
    net = UResNet(small_flags())
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(4, 5, dtype=torch.float64))


@pytest.mark.skipif(not os.path.exists("/root/reference/uresnet/models/uresnet_sparse.py"),
                    reason="reference tree only exists in the build container")
def test_reference_model_file_builds_unchanged_on_this_package():
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.model import UResNet
    scn.install()
    try:
        spec = importlib.util.spec_from_file_location("ref_uresnet_sparse",
                                                      "/root/reference/uresnet/models/uresnet_sparse.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        flags = small_flags(URESNET_FILTERS=16, URESNET_NUM_STRIDES=5, SPATIAL_SIZE=512)
        ref_net = mod.UResNet(flags)
        mine = UResNet(flags)
        assert list(ref_net.state_dict().keys()) == list(mine.state_dict().keys())
        assert [tuple(v.shape) for v in ref_net.state_dict().values()] == \
               [tuple(v.shape) for v in mine.state_dict().values()]
        assert isinstance(ref_net.sparseModel[0], scn.InputLayer) and ref_net.sparseModel[0].mode == 3
    finally:
        sys.modules.pop("sparseconvnet", None)
