"""Shared helpers of the parity tests."""
import numpy as np
import torch


def event_blob(n_events, ss, dim=3, seed=3, dup=0):
    """Synthetic blob (float64 [N, dim+2]) + labels; `dup` > 0 appends duplicated points
    (same coordinates, new values) to exercise InputLayer mode 3."""
    from uresnet_pytorch_b200 import synthetic
    data, label, counts = synthetic.make_batch(n_events, ss, dim=dim, seed=seed)
    if dup:
        rng = np.random.default_rng(seed + 99)
        pick = rng.integers(0, len(data), size=dup)
        extra = data[pick].copy()
        extra[:, -1] = rng.lognormal(-1.0, 0.8, size=dup)
        # some points three times
        data = np.concatenate([data, extra, extra[: dup // 3]], 0)
        label = np.concatenate([label, label[pick], label[pick][: dup // 3]], 0)
        perm = rng.permutation(len(data))
        data, label = data[perm], label[perm]
    return data, label


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).norm().item() / max(b.norm().item(), 1e-30)


def rules_as_set(rules):
    """oracle rule lists -> set of (k, in_row, out_row)"""
    s = set()
    for k, (ri, ro) in enumerate(rules):
        s.update((k, int(a), int(b)) for a, b in zip(ri, ro))
    return s


def table_as_set(table, n):
    """offset-major table [K, ld] -> set of (k, in_row, out_row) for rows < n"""
    t = table[:, :n].cpu().numpy()
    ks, outs = np.nonzero(t >= 0)
    return set((int(k), int(t[k, o]), int(o)) for k, o in zip(ks, outs))
