"""Block-overlap evaluation of the reference's S3DIS tester (/root/reference/s3dis_seg/evaluate_s3dis_with_overlap.py:
245-318), as host logic around any `predict_fn`.

A block holds more points than the network takes, so the tester draws `num_point` random points per block again and again,
adds every draw's logits onto the drawn points, and stops once every INNER point of every block of the batch has been drawn
at least once; a point's label is the argmax of its summed logits, and only inner points are scored.
`predict_fn(batch_input (b, num_point, D) float32) -> logits (b, num_point, num_classes)` is the network (e.g.
models.SPH3D_s3dis.get_model in inference mode); this module never touches the GPU itself.
"""
import numpy as np


def block_lengths(padded_all):
    """number of real rows per block of a padded batch (padding = -1 in the inner-label column)"""
    pad = padded_all[:, :, -1] < 0
    return np.where(pad.any(axis=1), pad.argmax(axis=1), padded_all.shape[1]).astype(np.int64)


def predict_blocks_with_overlap(padded_all, num_point, predict_fn, num_classes, rng=None, max_rounds=10000):
    """-> (list of (n_b, num_classes) float32 summed logits, list of (n_b,) int32 draw counts, rounds)"""
    rng = np.random.default_rng() if rng is None else rng
    b = padded_all.shape[0]
    lengths = block_lengths(padded_all)
    if (lengths == 0).any():
        raise ValueError("empty block in batch")
    summed = [np.zeros((int(n), num_classes), dtype=np.float32) for n in lengths]
    counts = [np.zeros((int(n),), dtype=np.int32) for n in lengths]
    inner = [padded_all[i, :int(n), -1] == 1 for i, n in enumerate(lengths)]
    batch_input = np.zeros((b, num_point, padded_all.shape[2] - 2), dtype=np.float32)
    rounds = 0
    while any((counts[i][inner[i]] == 0).any() for i in range(b)):
        if rounds >= max_rounds:
            raise RuntimeError("inner points still uncovered after %d rounds" % rounds)
        picks = []
        for i, n in enumerate(lengths):
            pick = rng.choice(int(n), num_point, replace=int(n) < num_point)
            picks.append(pick)
            batch_input[i] = padded_all[i, pick, 0:-2]
            np.add.at(counts[i], pick, 1)
        logits = np.asarray(predict_fn(batch_input), dtype=np.float32)
        if logits.shape != (b, num_point, num_classes):
            raise ValueError("predict_fn returned %s, expected %s" % (logits.shape, (b, num_point, num_classes)))
        for i in range(b):
            # the reference writes `pred_sum[sample_index] += pred`: with replacement a point drawn twice in one round
            # receives ONE of its logits rows (numpy fancy-index assignment keeps the last); same here
            summed[i][picks[i]] += logits[i]
        rounds += 1
    return summed, counts, rounds


class SegmentationMetrics(object):
    """running totals of evaluate_s3dis_with_overlap.py:306-316 over inner points: overall accuracy, per-class accuracy, IoU"""

    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.correct = 0
        self.seen = 0
        self.seen_class = np.zeros(num_classes, dtype=np.int64)
        self.correct_class = np.zeros(num_classes, dtype=np.int64)
        self.union_class = np.zeros(num_classes, dtype=np.int64)

    def update(self, summed_logits, gt_label, inner_label):
        pred = np.argmax(summed_logits, axis=1)
        keep = np.asarray(inner_label) == 1
        pred, gt = pred[keep], np.asarray(gt_label)[keep].astype(np.int64)
        self.correct += int((pred == gt).sum())
        self.seen += int(keep.sum())
        for c in range(self.num_classes):
            self.seen_class[c] += int((gt == c).sum())
            self.correct_class[c] += int(((pred == c) & (gt == c)).sum())
            self.union_class[c] += int(((pred == c) | (gt == c)).sum())

    def result(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            class_acc = self.correct_class / self.seen_class
            iou = self.correct_class / self.union_class
        return {"accuracy": self.correct / max(self.seen, 1), "class_accuracy": class_acc, "iou": iou,
                "mean_class_accuracy": float(np.nanmean(class_acc)), "mean_iou": float(np.nanmean(iou))}
