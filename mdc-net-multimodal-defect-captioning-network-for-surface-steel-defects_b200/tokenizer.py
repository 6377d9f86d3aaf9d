"""Token-id constants and the box/label part of the reference's token codec
(data_processing.py:224-290,317-391,556-598) -- the piece the hot path and its IoU stage need.

Sequence layout (data_processing.py:264-290):  BOS, <caption tokens>, label, xmin, ymin, xmax, ymax, EOS
with coordinate bins 0..num_bins-1, labels 258..267, UNK 299, BOS 300, EOS 301, PAD 302,
caption-start 303, caption-end 304, caption words from 270.  The spaCy vocabulary builder of the
reference is host-side data preparation and out of scope (`data_processing.py` row).
"""
from __future__ import annotations

import torch


class Tokenizer:
    UNK_code, BOS_code, EOS_code, PAD_code = 299, 300, 301, 302
    CAPTION_START, CAPTION_END = 303, 304
    LABEL_BASE = 258

    def __init__(self, num_classes=10, num_bins=224, width=224, height=224, max_len=100, vocab=None):
        self.num_classes, self.num_bins = num_classes, num_bins
        self.width, self.height, self.max_len = width, height, max_len
        self.vocab = vocab                      # optional {id: word}
        self.vocab_size = 305 if vocab is None else max(305, max(vocab) + 1)

    def quantize(self, x):
        return (x * (self.num_bins - 1)).round().long()

    def dequantize(self, x):
        return x.float() / (self.num_bins - 1)

    def decode_bboxes(self, token_batch):
        """data_processing.py:556-598 for a (B,L) batch: after the caption-end token, read groups of
        label,x0,y0,x1,y1; keep groups with 0<=v<=num_bins, x1>x0, y1>y0; de-quantise v/(num_bins-1)*W;
        zero-row pad to the longest.  Returns f32 (B,Nmax,4) (Nmax >= 1)."""
        out = []
        for seq in token_batch.tolist():
            boxes = []
            try:
                i = seq.index(self.CAPTION_END) + 1
            except ValueError:
                i = 1
            while i + 4 < len(seq):
                lab, x0, y0, x1, y1 = seq[i:i + 5]
                if lab in (self.EOS_code, self.PAD_code):
                    break
                if all(0 <= v <= self.num_bins for v in (x0, y0, x1, y1)) and x1 > x0 and y1 > y0:
                    s = 1.0 / (self.num_bins - 1)
                    boxes.append([x0 * s * self.width, y0 * s * self.height, x1 * s * self.width, y1 * s * self.height])
                i += 5
            out.append(boxes)
        n = max(1, max(len(b) for b in out))
        t = torch.zeros((len(out), n, 4), dtype=torch.float32)
        for b, boxes in enumerate(out):
            if boxes:
                t[b, :len(boxes)] = torch.tensor(boxes, dtype=torch.float32)
        return t

    def decode(self, tokens):
        """(labels, bboxes, caption words) of ONE sequence, data_processing.py:317-391 shape."""
        seq = tokens.tolist()
        caption = []
        if self.CAPTION_START in seq and self.CAPTION_END in seq:
            a, b = seq.index(self.CAPTION_START), seq.index(self.CAPTION_END)
            ids = seq[a + 1:b]
            caption = [self.vocab.get(t, "<unk>") if self.vocab else str(t) for t in ids]
        boxes = self.decode_bboxes(torch.tensor([seq]))[0]
        keep = boxes.abs().sum(-1) != 0
        labels = []
        try:
            i = seq.index(self.CAPTION_END) + 1
        except ValueError:
            i = 1
        while i + 4 < len(seq) and seq[i] not in (self.EOS_code, self.PAD_code):
            labels.append(seq[i]); i += 5
        return labels[:int(keep.sum())], boxes[keep], caption
