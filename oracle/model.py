"""ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).

The sparse U-ResNet graph of /root/reference/uresnet/models/uresnet_sparse.py:7-37 built on
the oracle's torch-CPU scn restatement (oracle/scn_cpu.py), plus the reference's
SegmentationLoss semantics (uresnet_sparse.py:40-82: per-event mean cross-entropy, summed
over events; accuracy averaged per event).
"""
import torch

from . import scn_cpu as scn


class OracleUResNet(torch.nn.Module):
    def __init__(self, flags):
        super().__init__()
        dim, m = flags.DATA_DIM, flags.URESNET_FILTERS
        planes = [i * m for i in range(1, flags.URESNET_NUM_STRIDES + 1)]  # uresnet_sparse.py:16
        self.sparseModel = (scn.Sequential()
                            .add(scn.InputLayer(dim, flags.SPATIAL_SIZE, mode=3))
                            .add(scn.SubmanifoldConvolution(dim, 1, m, 3, False))
                            .add(scn.UNet(dim, 2, planes, residual_blocks=True, downsample=[2, 2]))
                            .add(scn.BatchNormReLU(m))
                            .add(scn.OutputLayer(dim)))
        self.linear = torch.nn.Linear(m, flags.NUM_CLASS)

    def forward(self, point_cloud):
        dt = self.linear.weight.dtype
        coords = point_cloud[:, 0:-1].to(dt)          # uresnet_sparse.py:33
        features = point_cloud[:, -1][:, None].to(dt)  # uresnet_sparse.py:34
        x = self.sparseModel((coords, features))
        return [self.linear(x)]


def segmentation_loss(logits, data, label):
    """uresnet_sparse.py:57-82 for one replica: sum over events of the event-mean CE."""
    batch_ids = data[:, -2]
    ce = torch.nn.CrossEntropyLoss(reduction="none")
    total_loss, total_acc = 0.0, 0.0
    for b in batch_ids.unique():
        sel = batch_ids == b
        ev_logits = logits[sel]
        ev_label = label[sel].squeeze(-1).long()
        total_loss = total_loss + ce(ev_logits, ev_label).mean()
        total_acc += (ev_logits.argmax(-1) == ev_label).sum().item() / float(ev_label.nelement())
    return total_loss, total_acc
