"""Sparse U-ResNet and its loss on the b200 sparseconvnet surface.

Mirrors the behaviour of /root/reference/uresnet/models/uresnet_sparse.py: the same module
tree (hence the same state-dict keys: sparseModel.<idx>..., linear.*), the same
`forward(point_cloud[N, D+2]) -> [logits[N, nc]]` contract (:27-37) and the same loss
semantics (:40-82: per-event mean cross-entropy, summed over events; per-event accuracy).
The reference file itself also runs unchanged on this package; this
mirror exists because /root/reference is not present on the GPU box.
"""
import torch

from . import sparseconvnet as scn


class Flags(object):
    """Minimal stand-in for the reference's flags object (R/uresnet/flags.py): the models only
    read these attributes (uresnet_sparse.py:12-25)."""

    def __init__(self, DATA_DIM=3, SPATIAL_SIZE=512, URESNET_FILTERS=16, URESNET_NUM_STRIDES=5, NUM_CLASS=5,
                 BN_MOMENTUM=0.9, **kw):
        self.DATA_DIM, self.SPATIAL_SIZE = DATA_DIM, SPATIAL_SIZE
        self.URESNET_FILTERS, self.URESNET_NUM_STRIDES = URESNET_FILTERS, URESNET_NUM_STRIDES
        self.NUM_CLASS, self.BN_MOMENTUM = NUM_CLASS, BN_MOMENTUM
        for k, v in kw.items():
            setattr(self, k, v)


class UResNet(torch.nn.Module):
    def __init__(self, flags):
        super().__init__()
        self._flags = flags
        dim = flags.DATA_DIM
        width = flags.URESNET_FILTERS
        planes = [width * (lvl + 1) for lvl in range(flags.URESNET_NUM_STRIDES)]
        net = scn.Sequential()
        net.add(scn.InputLayer(dim, flags.SPATIAL_SIZE, mode=3))
        net.add(scn.SubmanifoldConvolution(dim, 1, width, 3, False))
        net.add(scn.UNet(dim, 2, planes, residual_blocks=True, downsample=[2, 2]))
        net.add(scn.BatchNormReLU(width))
        net.add(scn.OutputLayer(dim))
        self.sparseModel = net
        self.linear = torch.nn.Linear(width, flags.NUM_CLASS)

    def forward(self, point_cloud):
        # columns: D spatial coordinates, sample id, value (iotools_sparse.py:138-141,160)
        ncol = point_cloud.shape[1]
        coords = point_cloud[:, : ncol - 1]
        value = point_cloud[:, ncol - 1:].float()
        feats = self.sparseModel((coords, value))
        return [self.linear(feats)]


class SegmentationLoss(torch.nn.Module):
    """Sum over events of the event-mean cross-entropy (optionally pixel-weighted), and the
    sum of per-event accuracies -- the reference's semantics, computed without a Python loop
    over events or host syncs: per-event means via index_add on the sample-id column."""

    def __init__(self, flags=None, max_events=65535):
        super().__init__()
        self.max_events = max_events

    def forward(self, segmentation, data, label, weight=None):
        total_loss, total_acc = 0.0, 0.0
        for i in range(len(segmentation)):
            logits = segmentation[i]
            ids = data[i][:, -2].long()
            lab = label[i].reshape(-1).long()
            ce = torch.nn.functional.cross_entropy(logits, lab, reduction="none")
            if weight is not None:
                ce = ce * weight[i].reshape(-1).float()
            nb = self.num_events if getattr(self, "num_events", None) else int(ids.max().item()) + 1
            ones = torch.ones_like(ce)
            cnt = torch.zeros(nb, device=ce.device).index_add_(0, ids, ones)
            s = torch.zeros(nb, device=ce.device).index_add_(0, ids, ce)
            hit = (logits.argmax(-1) == lab).float()
            h = torch.zeros(nb, device=ce.device).index_add_(0, ids, hit)
            present = cnt > 0
            safe = torch.where(present, cnt, torch.ones_like(cnt))
            total_loss = total_loss + (s / safe).sum()
            total_acc = total_acc + (h / safe).sum()
        return total_loss, total_acc
