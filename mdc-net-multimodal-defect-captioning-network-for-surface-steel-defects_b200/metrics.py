"""Evaluation metrics of the reference's validation loop, evaluation-time only:

  MeanAveragePrecision   train_val_epoch.py:205-231 / :389-403 -- torchmetrics.detection.MeanAveragePrecision(box_format='xyxy',
                         iou_thresholds=[0.3]) with the same update([pred], [target]) / compute() calls.  The per-image greedy
                         prediction <-> ground-truth matching runs as ONE kernel for all collected images (csrc/metrics.cu,
                         mdc_map_match); the precision/recall accumulation over the matched set (a few thousand rows) is host-side
                         float64 numpy, following pycocotools COCOeval.accumulate (101 recall thresholds, area 'all', maxDets 100).
  calculate_bleu_scores  utils.py:56-63 -- nltk sentence_bleu with SmoothingFunction().method1, host-side.

torchmetrics, pycocotools and nltk are not installed here and not vendored by the reference: both functions are restated from
the published algorithms -- parity unpinned.
"""
from __future__ import annotations

import math
from collections import Counter

import numpy as np
import torch

from . import _lib as L
from .config import CFG


class MeanAveragePrecision:
    def __init__(self, box_format="xyxy", iou_thresholds=None, class_metrics=False, max_detection_threshold=100):
        if box_format != "xyxy":
            raise ValueError("only box_format='xyxy' (what the reference passes, train_val_epoch.py:210) is implemented")
        self.iou_thresholds = list(iou_thresholds) if iou_thresholds is not None else [0.5 + 0.05 * i for i in range(10)]
        self.class_metrics, self.max_det = bool(class_metrics), int(max_detection_threshold)
        self.device = None
        self.preds, self.targets = [], []

    def to(self, device):
        self.device = torch.device(device)
        return self

    def update(self, preds, targets):
        if len(preds) != len(targets):
            raise ValueError("preds and targets must have the same length")
        for p, t in zip(preds, targets):
            self.preds.append({k: p[k].detach() for k in ("boxes", "scores", "labels")})
            self.targets.append({k: t[k].detach() for k in ("boxes", "labels")})

    def reset(self):
        self.preds, self.targets = [], []

    def _match(self, thr):
        """-> per image numpy arrays (scores, labels, match) of the evaluated detections + ground-truth labels per image."""
        dev = self.device or (CFG.device if CFG.device.type == "cuda" else None)
        if dev is None or torch.device(dev).type != "cuda":
            raise L.MdcError("mAP matching runs on the GPU (mdc_map_match); there is no CPU fallback")
        dev = torch.device(dev)
        B = len(self.preds)
        N = max(1, max(int(p["boxes"].shape[0]) for p in self.preds))
        M = max(1, max(int(t["boxes"].shape[0]) for t in self.targets))
        if N > 128 or M > 128:
            raise ValueError("mdc_map_match handles at most 128 detections and 128 ground-truth boxes per image")
        pb = torch.zeros((B, N, 4), dtype=torch.float32); ps = torch.zeros((B, N), dtype=torch.float32); pl = torch.zeros((B, N), dtype=torch.int32)
        gb = torch.zeros((B, M, 4), dtype=torch.float32); gl = torch.zeros((B, M), dtype=torch.int32)
        npred = torch.zeros(B, dtype=torch.int32); ngt = torch.zeros(B, dtype=torch.int32)
        for i, (p, t) in enumerate(zip(self.preds, self.targets)):
            n, m = int(p["boxes"].shape[0]), int(t["boxes"].shape[0])
            npred[i], ngt[i] = n, m
            if n:
                pb[i, :n] = p["boxes"].reshape(n, 4).float().cpu(); ps[i, :n] = p["scores"].float().cpu(); pl[i, :n] = p["labels"].to(torch.int32).cpu()
            if m:
                gb[i, :m] = t["boxes"].reshape(m, 4).float().cpu(); gl[i, :m] = t["labels"].to(torch.int32).cpu()
        d = [x.to(dev).contiguous() for x in (pb, ps, pl, npred, gb, gl, ngt)]
        match = torch.empty((B, N), dtype=torch.int32, device=dev); order = torch.empty((B, N), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().mdc_map_match(L.ctx(dev), L.ptr(d[0]), L.ptr(d[1]), L.ptr(d[2]), L.ptr(d[3]), L.ptr(d[4]), L.ptr(d[5]), L.ptr(d[6]),
                                          B, N, M, float(thr), self.max_det, L.ptr(match), L.ptr(order), L.stream_ptr(dev)))
        return ps.numpy(), pl.numpy(), match.cpu().numpy(), order.cpu().numpy(), gl.numpy(), ngt.numpy()

    def compute(self):
        """Returns the torchmetrics result keys the reference reads: 'map' (mean over the IoU thresholds and the classes that have
        ground truth) and, with class_metrics, 'map_per_class' / 'classes'.  -1 where no ground truth exists at all."""
        if not self.preds:
            return {"map": torch.tensor(-1.0)}
        rec_thrs = np.linspace(0.0, 1.0, 101)
        per_thr_class = []
        classes = None
        for thr in self.iou_thresholds:
            scores, labels, match, order, gl, ngt = self._match(thr)
            cls_all = sorted(set(int(c) for i in range(len(ngt)) for c in gl[i, :ngt[i]]) | set(int(c) for c in labels[match > -2]))
            classes = cls_all
            aps = []
            for c in cls_all:
                npig = int(sum((gl[i, :ngt[i]] == c).sum() for i in range(len(ngt))))
                if npig == 0:
                    aps.append(-1.0)
                    continue
                # detections of class c over all images, per image in score order (the kernel's order), then a STABLE global sort
                sc, tp = [], []
                for i in range(scores.shape[0]):
                    idx = np.nonzero((labels[i] == c) & (match[i] > -2))[0]
                    idx = idx[np.argsort(order[i, idx], kind="mergesort")]
                    sc.append(scores[i, idx]); tp.append(match[i, idx] >= 0)
                sc = np.concatenate(sc) if sc else np.zeros(0, np.float32); tp = np.concatenate(tp) if tp else np.zeros(0, bool)
                inds = np.argsort(-sc, kind="mergesort")
                tps = np.cumsum(tp[inds].astype(np.float64)); fps = np.cumsum((~tp[inds]).astype(np.float64))
                rc = tps / npig
                pr = tps / (fps + tps + np.spacing(1))
                q = np.zeros(101)
                pr = pr.tolist()
                for i in range(len(pr) - 1, 0, -1):
                    if pr[i] > pr[i - 1]:
                        pr[i - 1] = pr[i]
                ii = np.searchsorted(rc, rec_thrs, side="left")
                for ri, pi in enumerate(ii):
                    if pi < len(pr):
                        q[ri] = pr[pi]
                aps.append(float(q.mean()))
            per_thr_class.append(aps)
        arr = np.array(per_thr_class, dtype=np.float64)           # (thresholds, classes)
        valid = arr > -1
        out = {"map": torch.tensor(float(arr[valid].mean()) if valid.any() else -1.0)}
        if self.class_metrics:
            pc = [float(arr[:, j][valid[:, j]].mean()) if valid[:, j].any() else -1.0 for j in range(arr.shape[1])]
            out["map_per_class"] = torch.tensor(pc)
            out["classes"] = torch.tensor(classes, dtype=torch.int32)
        return out


def _modified_precision(reference, hypothesis, n):
    ref_counts = Counter(tuple(reference[i:i + n]) for i in range(len(reference) - n + 1)) if len(reference) >= n else Counter()
    hyp_counts = Counter(tuple(hypothesis[i:i + n]) for i in range(len(hypothesis) - n + 1)) if len(hypothesis) >= n else Counter()
    clipped = {g: min(c, ref_counts.get(g, 0)) for g, c in hyp_counts.items()}
    return sum(clipped.values()), max(1, sum(hyp_counts.values()))


def sentence_bleu_method1(reference, hypothesis, weights=(0.25, 0.25, 0.25, 0.25), epsilon=0.1):
    """nltk.translate.bleu_score.sentence_bleu([reference], hypothesis, smoothing_function=SmoothingFunction().method1):
    modified n-gram precisions with clipping, zero numerators replaced by epsilon (method1), brevity penalty, geometric mean."""
    p = [_modified_precision(reference, hypothesis, n) for n in range(1, len(weights) + 1)]
    if p[0][0] == 0:
        return 0.0
    hyp_len, ref_len = len(hypothesis), len(reference)
    bp = 1.0 if hyp_len > ref_len else (0.0 if hyp_len == 0 else math.exp(1 - ref_len / hyp_len))
    s = 0.0
    for w, (num, den) in zip(weights, p):
        s += w * math.log((num if num != 0 else epsilon) / den)
    return bp * math.exp(s)


def calculate_bleu_scores(ground_truths, predictions):
    """utils.py:56-63: one smoothed sentence-BLEU per (reference token list, predicted token list) pair."""
    return [sentence_bleu_method1(ref, pred) for ref, pred in zip(ground_truths, predictions)]
