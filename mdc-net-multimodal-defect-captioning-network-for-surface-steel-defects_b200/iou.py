"""Box-IoU scoring with the reference's function names, argument order and return containers
(iou_calcualtions.py:5-255, iou_bbox.py:3-63) on one batched CUDA kernel instead of a per-image Python
loop of ~12 broadcast launches and a `.tolist()` sync per image.

Inputs are xyxy boxes; zero rows are padding (data_processing.py:596 pad_sequence) and are scored like
the reference scores them (Q13).  Everything is float32 (the reference's dtype for decoded boxes).
"""
from __future__ import annotations

import torch

from . import _lib as L
from .config import CFG


def _dev(*ts):
    for t in ts:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if CFG.device.type == "cuda":
        return CFG.device
    raise L.MdcError("IoU scoring needs a CUDA device; there is no CPU fallback")


def _prep(x, dev):
    return x.to(dev, torch.float32).contiguous()


def _batched(mode, pred, gt, want_iou=True, want_max=False):
    """pred (B,N,4), gt (B,M,4) -> (iou (B,N,M) | None, max (B,N) | None) on device."""
    dev = _dev(pred, gt)
    pred, gt = _prep(pred, dev), _prep(gt, dev)
    B, N, _ = pred.shape
    M = gt.shape[1]
    iou = torch.empty((B, N, M), dtype=torch.float32, device=dev) if want_iou else None
    mx = torch.empty((B, N), dtype=torch.float32, device=dev) if want_max else None
    if B * N * M > 0:
        with torch.cuda.device(dev):
            L.check(L.lib().mdc_iou_batch(L.ctx(dev), mode, L.ptr(pred), L.ptr(gt), B, N, M, L.ptr(iou), L.ptr(mx), L.stream_ptr()))
    return iou, mx


def bbox_iou(box1, box2):
    """iou_calcualtions.py:5-40: (N,4),(M,4) -> (N,M), inter / (union + 1e-6)."""
    return _batched(L.IOU_EPS, box1[None], box2[None])[0][0]


def calculate_batch_iou(predicted_bboxes, ground_truth_bboxes):
    """iou_calcualtions.py:45-56: returns a list of B tensors (N,M)."""
    iou, _ = _batched(L.IOU_EPS, predicted_bboxes, ground_truth_bboxes)
    return list(iou.unbind(0))


def calculate_batch_max_iou(predicted_bboxes, ground_truth_bboxes):
    """iou_calcualtions.py:59-75: flat python list of B*N row-maxima (zero-padded rows included)."""
    if predicted_bboxes.size(1) == 0 or ground_truth_bboxes.size(1) == 0:
        return []
    _, mx = _batched(L.IOU_EPS, predicted_bboxes, ground_truth_bboxes, want_iou=False, want_max=True)
    return mx.flatten().tolist()       # ONE device->host copy for the whole batch


def calculate_batch_max_iou_torchvision(predicted_bboxes, ground_truth_bboxes):
    """iou_calcualtions.py:78-105: torchvision.ops.box_iou (no epsilon) + nan_to_num(0) + row max."""
    if predicted_bboxes.dim() == 2:
        predicted_bboxes = predicted_bboxes[:, None, :]
    if ground_truth_bboxes.dim() == 2:
        ground_truth_bboxes = ground_truth_bboxes[:, None, :]
    if predicted_bboxes.size(1) == 0 or ground_truth_bboxes.size(1) == 0:
        return []
    _, mx = _batched(L.IOU_NAN0, predicted_bboxes, ground_truth_bboxes, want_iou=False, want_max=True)
    return mx.flatten().tolist()


def calculate_batch_max_iou_masked(predicted_bboxes, ground_truth_bboxes):
    """Variant offered next to the reference behaviour (Q13): maxima of zero-padded prediction rows are
    dropped.  Returns a flat python list like calculate_batch_max_iou."""
    _, mx = _batched(L.IOU_EPS, predicted_bboxes, ground_truth_bboxes, want_iou=False, want_max=True)
    keep = predicted_bboxes.to(mx.device).abs().sum(-1) != 0
    return mx[keep].tolist()


def giou_pairwise(pred_boxes, gt_boxes):
    """iou_calcualtions.py:220-255: (N,4),(M,4) -> (N,M)."""
    return _batched(L.IOU_GIOU, pred_boxes[None], gt_boxes[None])[0][0]


def giou_loss_with_scores(pred_boxes, gt_boxes, no_detection_penalty=1.0):
    """iou_calcualtions.py:165-208: (total loss scalar tensor, list of per-image GIoU matrices with the
    all-zero rows/columns filtered out)."""
    dev = _dev(pred_boxes, gt_boxes)
    pred, gt = _prep(pred_boxes, dev), _prep(gt_boxes, dev)
    B, N, _ = pred.shape
    M = gt.shape[1]
    loss = torch.empty(B + 1, dtype=torch.float32, device=dev)
    giou = torch.empty((B, N, M), dtype=torch.float32, device=dev)
    valid = torch.empty((B, N, M), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().mdc_giou_loss(L.ctx(dev), L.ptr(pred), L.ptr(gt), B, N, M, float(no_detection_penalty),
                                      L.ptr(loss), L.ptr(giou), L.ptr(valid), L.stream_ptr()))
    scores = []
    vp = valid.bool().any(dim=2)      # (B,N) rows with a non-zero-sum prediction and at least one valid GT
    vg = valid.bool().any(dim=1)      # (B,M)
    for i in range(B):                # container shaping only (ragged python list like the reference)
        if vp[i].any() and vg[i].any():
            scores.append(giou[i][vp[i]][:, vg[i]])
        else:
            scores.append(torch.zeros(0, device=dev))
    return loss[B], scores


def calculate_iou(pred_boxes, gt_boxes):
    """iou_bbox.py:3-43: inter / union without epsilon (0/0 -> nan)."""
    if pred_boxes.nelement() == 0 or gt_boxes.nelement() == 0:
        return torch.tensor(0.0)
    if pred_boxes.dim() < 2:
        pred_boxes = pred_boxes.unsqueeze(0)
    if gt_boxes.dim() < 2:
        gt_boxes = gt_boxes.unsqueeze(0)
    return _batched(L.IOU_PLAIN, pred_boxes[None], gt_boxes[None])[0][0]


def iou_loss(pred_boxes, gt_boxes, min_penalty=0.5):
    """iou_bbox.py:46-63: mean(1 - iou) with zero / nan IoUs replaced by `min_penalty`."""
    ious = calculate_iou(pred_boxes, gt_boxes)
    ious = torch.where(ious > 0, ious, torch.tensor(min_penalty, device=ious.device))
    return (1 - ious).mean()
