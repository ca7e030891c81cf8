import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from conftest import small_flags
from helpers import event_blob, rel_l2
from uresnet_pytorch_b200.trainer import Trainer
flags = small_flags()
blobs = [event_blob(2, 32, seed=s) for s in (1, 2, 3, 4)]
cap = max(len(b[0]) for b in blobs) + 37
for ov_e, ov_g in [(False, False), (True, False), (False, True), (True, True)]:
    te = Trainer(flags, lr=1e-3, graph=False, events_per_step=2, seed=5, overlap_wgrad=ov_e)
    tg = Trainer(flags, lr=1e-3, graph=True, capacity=cap, events_per_step=2, seed=5, overlap_wgrad=ov_g)
    out = []
    for data, label in blobs:
        a = te.train_step(data, label); b = tg.train_step(data, label)
        out.append((round(a['loss'], 5), round(b['loss'], 5)))
    worst = max(rel_l2(b, a) for (n, a), (_, b) in zip(te.net.named_parameters(), tg.net.named_parameters()))
    print('overlap eager/graph', ov_e, ov_g, out, 'worst param rel', worst)
