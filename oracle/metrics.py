"""ORACLE -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the accuracy / confusion block of the reference's per-event inference metrics,
/root/reference/uresnet/utils.py:127-200 (`compute_metrics_sparse`): for every event of the batch-id column
(:131-136) the non-zero accuracy (:138-139), the log-loss (:141-142; sklearn.metrics.log_loss restated as the mean of
-log softmax[label]), the mean softmax of the true class (:145-146), pixel counts (:148), and per class: accuracy
(:188), mean softmax of the true class (:191), pixel count (:192-195), confusion and energy-confusion rows
(:196-199).  The DBSCAN / Michel / distance-histogram parts are analysis scripts outside the hot path.

Pinned against the reference file itself where it can run (tests/test_oracle.py imports utils.py by path when
/root/reference, sklearn and scipy are present and compares on random inputs).
"""
import numpy as np


def compute_metrics_sparse(data_v, label_v, softmax_v):
    res = {k: [] for k in ("acc", "correct_softmax", "id", "nonzero_pixels", "class_acc", "class_pixel",
                           "class_mean_softmax", "confusion_matrix", "energy_confusion_matrix", "loss_seg")}
    for i, label in enumerate(label_v):
        data, softmax = data_v[i], softmax_v[i]
        for batch_id in np.unique(data[:, -2]):                                    # utils.py:131
            sel = data[:, -2] == batch_id
            ev_data, ev_sm, ev_lab = data[sel], softmax[sel], label[sel]
            pred = np.argmax(ev_sm, axis=1)[:, None]                               # :138
            res["acc"].append((ev_lab == pred).astype(np.int32).sum() / float(len(ev_lab)))
            lab_i = ev_lab.reshape(-1).astype(np.int32)
            p = np.clip(ev_sm[np.arange(len(lab_i)), lab_i], 1e-15, 1.0)
            res["loss_seg"].append(float(-np.log(p).mean()))                        # :141 (log_loss)
            correct = ev_sm[np.arange(len(lab_i)), lab_i][:, None]                  # :145
            res["correct_softmax"].append(np.mean(correct))
            res["id"].append(batch_id)
            res["nonzero_pixels"].append(ev_lab.shape[0])
            nc = ev_sm.shape[1]
            conf = np.zeros((nc, nc), np.int32)
            econf = np.zeros((nc, nc), np.float32)
            class_acc, class_pix, class_sm = [], [], []
            for c in range(nc):                                                    # :186
                ci = ev_lab == c
                with np.errstate(divide="ignore", invalid="ignore"):
                    class_acc.append((ev_lab[ci] == pred[ci]).astype(np.int32).sum() / np.float64(len(ev_lab[ci])))
                    class_sm.append(np.mean(correct[ci]) if ci.any() else np.nan)
                class_pix.append(int(ci.sum()))
                for c2 in range(nc):
                    cj = pred[ci] == c2
                    conf[c][c2] = cj.astype(np.int32).sum()
                    econf[c][c2] = ev_data[..., -1][..., None][ci][cj].sum()
            res["class_acc"].append(np.array(class_acc))
            res["class_pixel"].append(np.array(class_pix))
            res["class_mean_softmax"].append(np.array(class_sm))
            res["confusion_matrix"].append(conf)
            res["energy_confusion_matrix"].append(econf)
    return res
