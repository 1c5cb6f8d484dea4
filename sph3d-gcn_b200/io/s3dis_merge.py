"""Scene merge of the block predictions -- the numpy restatement of /root/reference/post-merging/s3dis_merge.m:37-95
(there is no MATLAB / octave in this image, scipy's k-d tree stands in for knnsearch).

The tester (io/s3dis_eval.py = s3dis_seg/evaluate_s3dis_with_overlap.py) leaves, per block, the summed logits of every
block point, its inner flag and `index_label` = the row of the (3 cm voxelised) scene cloud the point came from.  The
merge then, per scene:
  1. keeps the INNER points of each block (s3dis_merge.m:42-44);
  2. turns their summed logits into a confidence: L2-normalise the logit vector, then softmax it (:45-46);
  3. accumulates the confidences of all blocks onto the scene points through `index_label` (:48, :56);
  4. labels every scene point with the argmax (:59-60; first maximum, as MATLAB's max);
  5. hands each point of the FULL-resolution cloud the label of its nearest voxelised point (:73-76);
  6. accumulates per-class intersection / union / seen over the full cloud (:77-81).
Host-side numpy only; nothing here touches the GPU.
"""
import numpy as np


def block_confidence(summed_logits):
    """(n, C) summed logits -> (n, C) confidences: unit-normalise each row, then softmax it (s3dis_merge.m:45-46).
    A row of zeros (a point never drawn) would be 0/0 in the MATLAB code and poison its scene point with NaN; here it
    contributes nothing instead (it cannot occur after predict_blocks_with_overlap: every inner point is covered)."""
    x = np.asarray(summed_logits, dtype=np.float64)
    norm = np.sqrt((x * x).sum(axis=1, keepdims=True))
    ok = norm[:, 0] > 0
    unit = np.zeros_like(x)
    unit[ok] = x[ok] / norm[ok]
    e = np.exp(unit)
    conf = e / e.sum(axis=1, keepdims=True)
    conf[~ok] = 0.0
    return conf


def merge_scene(num_scene_points, blocks, num_classes):
    """blocks: iterable of (summed_logits (n, C), inner_label (n,), index_label (n,) 0-based rows of the scene cloud).
    -> (predictions (P, C) float64 accumulated confidences, pred_label (P,) int64)"""
    predictions = np.zeros((int(num_scene_points), int(num_classes)), dtype=np.float64)
    for summed_logits, inner_label, index_label in blocks:
        keep = np.asarray(inner_label).reshape(-1) == 1
        idx = np.asarray(index_label).reshape(-1)[keep].astype(np.int64)
        if idx.size and (idx.min() < 0 or idx.max() >= num_scene_points):
            raise ValueError("index_label outside the scene cloud")
        conf = block_confidence(np.asarray(summed_logits)[keep])
        # MATLAB's predictions(idx,:) = predictions(idx,:) + conf: a row index repeated INSIDE one block keeps its last
        # write, exactly like numpy's fancy-index assignment (the inner points of a block are distinct scene points)
        predictions[idx] = predictions[idx] + conf
    return predictions, np.argmax(predictions, axis=1)


def propagate_to_full_cloud(voxel_xyz, voxel_label, full_xyz):
    """label of the nearest voxelised point for every point of the full-resolution cloud (knnsearch, s3dis_merge.m:73-76)"""
    from scipy.spatial import cKDTree
    _, nearest = cKDTree(np.asarray(voxel_xyz, dtype=np.float64)).query(np.asarray(full_xyz, dtype=np.float64), k=1)
    return np.asarray(voxel_label)[nearest]


class SceneIoU(object):
    """total_intersect / total_union / total_seen of s3dis_merge.m:17-19, :77-81, accumulated over scenes"""

    def __init__(self, num_classes):
        self.num_classes = int(num_classes)
        self.intersect = np.zeros(self.num_classes, dtype=np.int64)
        self.union = np.zeros(self.num_classes, dtype=np.int64)
        self.seen = np.zeros(self.num_classes, dtype=np.int64)

    def update(self, pred_label, gt_label):
        pred, gt = np.asarray(pred_label).reshape(-1), np.asarray(gt_label).reshape(-1)
        for c in range(self.num_classes):
            self.intersect[c] += int(((pred == c) & (gt == c)).sum())
            self.union[c] += int(((pred == c) | (gt == c)).sum())
            self.seen[c] += int((gt == c).sum())

    def result(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = self.intersect / self.union
            acc = self.intersect / self.seen
        return {"iou": iou, "mean_iou": float(np.nanmean(iou)), "class_accuracy": acc,
                "overall_accuracy": float(self.intersect.sum() / max(int(self.seen.sum()), 1))}
