"""ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).

"Oracle B": an arithmetic cross-check that does not depend on any recollection of scn.
On a small grid the sparse layers are *defined* by their dense counterparts restricted to
the active sites (inactive sites hold zeros and no layer on the path has a bias,
/root/reference/uresnet/models/uresnet_sparse.py:21-22 pass bias=False):

    SubmanifoldConvolution k=3  ==  F.conv{2,3}d(dense, W_t, padding=1)   sampled at active sites
    Convolution k=2 s=2         ==  F.conv{2,3}d(dense, W_t, stride=2)     sampled at coarse sites
    Deconvolution k=2 s=2       ==  F.conv_transpose{2,3}d(dense, W_t, stride=2) at fine sites
    BatchNormReLU               ==  relu(BatchNorm1d(C, eps=1e-4, momentum=0.1)(features))

with W_t[co, ci, kx, ky, kz] = W[k, ci, co], k enumerating the filter box with the last
dimension fastest.
"""
import torch
import torch.nn.functional as F


def densify(coords, feats, ss, n_batch):
    """coords int64 [n, D+1] (last = sample) -> dense [B, C, ss, (ss,) ss]."""
    D = coords.shape[1] - 1
    dense = torch.zeros((n_batch, feats.shape[1]) + (ss,) * D, dtype=feats.dtype)
    dense[coords[:, D], :, *[coords[:, d] for d in range(D)]] = feats
    return dense


def sample(dense, coords):
    D = coords.shape[1] - 1
    return dense[coords[:, D], :, *[coords[:, d] for d in range(D)]]


def _wt(W, k, D):
    # W [k^D, nIn, nOut] -> [nOut, nIn, k, .., k]
    return W.permute(2, 1, 0).reshape((W.shape[2], W.shape[1]) + (k,) * D).contiguous()


def subm_dense(coords, feats, W, ss, n_batch, k=3):
    D = coords.shape[1] - 1
    conv = F.conv3d if D == 3 else F.conv2d
    out = conv(densify(coords, feats, ss, n_batch), _wt(W, k, D), padding=k // 2)
    return sample(out, coords)


def down_dense(coords, feats, W, ss, n_batch, coarse_coords, s=2):
    D = coords.shape[1] - 1
    conv = F.conv3d if D == 3 else F.conv2d
    out = conv(densify(coords, feats, ss, n_batch), _wt(W, s, D), stride=s)
    return sample(out, coarse_coords)


def up_dense(coarse_coords, feats, W, ss_coarse, n_batch, fine_coords, s=2):
    D = fine_coords.shape[1] - 1
    convt = F.conv_transpose3d if D == 3 else F.conv_transpose2d
    # conv_transpose weight is [in, out, k..]; W is [k^D, nIn, nOut]
    wt = W.permute(1, 2, 0).reshape((W.shape[1], W.shape[2]) + (s,) * D).contiguous()
    out = convt(densify(coarse_coords, feats, ss_coarse, n_batch), wt, stride=s)
    return sample(out, fine_coords)


def bnrelu_dense(feats, weight, bias, eps=1e-4):
    bn = torch.nn.BatchNorm1d(feats.shape[1], eps=eps, momentum=0.1).to(feats.dtype)
    with torch.no_grad():
        bn.weight.copy_(weight)
        bn.bias.copy_(bias)
    bn.train()
    return F.relu(bn(feats)), bn.running_mean.clone(), bn.running_var.clone()
