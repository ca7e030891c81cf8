"""ORACLE -- TEST INFRASTRUCTURE ONLY.

The dense path has a *live* reference: /root/reference/uresnet/models/uresnet_dense.py is plain
torch and imports under Python 3 once `xrange` exists.  This module loads that
file BY PATH (nothing is copied into the repo) so tests/golden/make_dense_golden.py can run the
unmodified reference model on torch-CPU and commit its input/output vectors as fixtures.
/root/reference only exists in the build container; on the GPU box the fixtures are used.
"""
import builtins
import importlib.util
import os

REF_DENSE = "/root/reference/uresnet/models/uresnet_dense.py"


def available():
    return os.path.exists(REF_DENSE)


def build():
    """Nothing to compile: the dense reference is Python.  Kept so __graft_entry__.build() has one
    hook for 'prepare the reference-backed checker'."""
    return available()


def load():
    """Import the reference's dense model file unchanged; returns the module."""
    if not available():
        raise RuntimeError("reference tree not present")
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range  # the file is Python 2 (uresnet_dense.py:141,155,212,216)
    spec = importlib.util.spec_from_file_location("ref_uresnet_dense", REF_DENSE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
