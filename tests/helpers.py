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


def check_grads(net, ref, global_tol, tensor_tol):
    """Gradient parity of two models with identical parameter names.

    Deep BN+ReLU stacks amplify round-off (a 1e-7 perturbation can flip a ReLU mask), and
    tensors whose gradient nearly cancels have a large *relative* error for a tiny absolute
    one; so the primary bound is the relative L2 error of the whole flattened gradient, with
    a looser per-tensor bound as a guard against a single broken layer."""
    import torch
    g_ref = dict(ref.named_parameters())
    num = den = 0.0
    for name, p in net.named_parameters():
        a, b = p.grad.detach().double().cpu(), g_ref[name].grad.detach().double()
        num += (a - b).pow(2).sum().item()
        den += b.pow(2).sum().item()
        assert rel_l2(a, b) <= tensor_tol, "%s: rel-L2 %g" % (name, rel_l2(a, b))
    glob = (num / max(den, 1e-300)) ** 0.5
    assert glob <= global_tol, "global gradient rel-L2 %g" % glob
    return glob
