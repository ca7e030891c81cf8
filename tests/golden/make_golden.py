"""Generate tests/golden/sparse_tiny.npz from the CPU oracle (oracle/).

The reference ships no golden vectors for the sparse path and its arithmetic lives in the
un-vendored `sparseconvnet`; the fixture therefore pins the ORACLE (parity unpinned against
scn itself).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import event_blob  # noqa: E402
from oracle.model import OracleUResNet, segmentation_loss  # noqa: E402
from uresnet_pytorch_b200.model import Flags  # noqa: E402


def main():
    ss, uf, uns = 32, 8, 3
    torch.manual_seed(20261018)
    m = OracleUResNet(Flags(SPATIAL_SIZE=ss, URESNET_FILTERS=uf, URESNET_NUM_STRIDES=uns)).train()
    with torch.no_grad():  # non-trivial BN affine parameters
        for mod in m.modules():
            if hasattr(mod, "running_mean"):
                mod.weight.uniform_(0.7, 1.3)
                mod.bias.uniform_(-0.2, 0.2)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    data, label = event_blob(2, ss, dup=30, seed=11)
    d, l = torch.from_numpy(data), torch.from_numpy(label)
    out = m(d)[0]
    loss, acc = segmentation_loss(out, d, l)
    loss.backward()
    blob = {"ss": ss, "uf": uf, "uns": uns, "data": data, "label": label,
            "logits": out.detach().numpy(), "loss": loss.item(), "acc": acc,
            "grad_stem": m.sparseModel[1].weight.grad.numpy(),
            "grad_linear": m.linear.weight.grad.numpy()}
    for k, v in sd0.items():
        blob["sd." + k] = v.numpy()
    for k, v in m.state_dict().items():
        if "running" in k:
            blob["after." + k] = v.numpy()
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sparse_tiny.npz")
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, os.path.getsize(out_path), "bytes; N =", len(data), "loss =", loss.item())


if __name__ == "__main__":
    main()
