"""TEST INFRASTRUCTURE ONLY -- CPU restatement of MDC-Net's batched inference hot path.

This file is the parity ORACLE for the CUDA path.  It is written as explicit tensor algebra
over a flat state-dict (the reference's own key names) so that it
 (a) travels to the GPU box (where /root/reference does not exist),
 (b) can run in float64 to arbitrate between two fp32 implementations, and
 (c) is pinned: oracle/make_golden.py runs the UNMODIFIED reference files
     (/root/reference/model.py, axial_model.py, iou_calcualtions.py, iou_bbox.py, through
     the import shims in oracle/shims) on the same weights/inputs, asserts agreement with
     this restatement, and commits the reference's outputs under tests/golden/.
     The ViT backbone (third-party `timm`, absent) is PARITY UNPINNED -- see oracle/shims/timm.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this module.  The product package never does; it fails loudly without its CUDA library.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# configuration (mirrors the globals the reference reads: allied_files.py:15-36,
# model.py:32,60,94,117 and utils.py:8,29)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleCfg:
    max_len: int = 100          # allied_files.py:18
    pad_idx: int = 302          # data_processing.py Tokenizer.PAD_code (token ids)
    bos_idx: int = 300
    enc_heads: int = 8          # deit3_medium: 8 heads x 64
    dec_heads: int = 8          # inference_p.py:128
    out_dim: int = 256          # inference_p.py:126
    axial_heads: int = 8        # axial_model.py:20
    axial_scale: float = 64 ** -0.5   # axial_model.py:23 (default dim_head, hard-wired)


def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


# ----------------------------------------------------------------------------------------
# A.1 encoder: timm VisionTransformer (DeiT-III) + channel pooling    model.py:14-23
# ----------------------------------------------------------------------------------------
def patches(x, p=16):
    """(B,C,H,W) -> (B, n, C*p*p); a patch flattens in (ch, dy, dx) order, patches row-major.
    Equivalent to Conv2d(k=p, s=p) as a matmul (timm PatchEmbed.proj)."""
    B, C, H, W = x.shape
    x = x.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, (H // p) * (W // p), C * p * p)


def vit_tokens(sd, x, heads, prefix="encoder.model.", eps=1e-6, upto=None):
    """All 1+n tokens after the final norm (global_pool='', num_classes=0)."""
    g = lambda k: sd[prefix + k]
    D = g("pos_embed").shape[-1]
    h = patches(x) @ g("patch_embed.proj.weight").reshape(D, -1).T + g("patch_embed.proj.bias")
    h = h + g("pos_embed")                                    # no_embed_class: before cls
    h = torch.cat([g("cls_token").expand(h.shape[0], -1, -1), h], dim=1)
    B, N, _ = h.shape
    hd = D // heads
    depth = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in sd if k.startswith(prefix + "blocks."))
    if upto is not None:
        depth = min(depth, upto)
    for i in range(depth):
        b = lambda k: g(f"blocks.{i}.{k}")
        u = _ln(h, b("norm1.weight"), b("norm1.bias"), eps)
        qkv = u @ b("attn.qkv.weight").T + b("attn.qkv.bias")
        q, k, v = [t.reshape(B, N, heads, hd).transpose(1, 2) for t in qkv.split(D, dim=-1)]
        a = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2), dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, N, D)
        h = h + b("ls1.gamma") * (a @ b("attn.proj.weight").T + b("attn.proj.bias"))
        u = _ln(h, b("norm2.weight"), b("norm2.bias"), eps)
        m = _gelu_erf(u @ b("mlp.fc1.weight").T + b("mlp.fc1.bias"))
        h = h + b("ls2.gamma") * (m @ b("mlp.fc2.weight").T + b("mlp.fc2.bias"))
    if upto is not None:
        return h
    return _ln(h, g("norm.weight"), g("norm.bias"), eps)


def adaptive_avg_pool_channels(x, out_dim):
    """nn.AdaptiveAvgPool1d over the LAST axis (model.py:19,23): bin i averages
    [floor(i*C/out), ceil((i+1)*C/out))."""
    C = x.shape[-1]
    cols = []
    for i in range(out_dim):
        s = (i * C) // out_dim
        e = -((-(i + 1) * C) // out_dim)
        cols.append(x[..., s:e].mean(-1))
    return torch.stack(cols, dim=-1)


def encoder_forward(sd, x, cfg: OracleCfg):
    """Encoder.forward, model.py:21-23: features[:, 1:] -> AdaptiveAvgPool1d(out_dim)."""
    feats = vit_tokens(sd, x, cfg.enc_heads)
    return adaptive_avg_pool_channels(feats[:, 1:], cfg.out_dim)


# ----------------------------------------------------------------------------------------
# A.2/A.3 decoder: torch.nn.TransformerDecoder post-norm stack, restated
# (torch/nn/modules/transformer.py:1158-1197 _sa_block/_mha_block/_ff_block,
#  torch/nn/functional.py multi_head_attention_forward; float masks are ADDED, Q7)
# ----------------------------------------------------------------------------------------
def _num_dec_layers(sd, prefix):
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix))


def _mha(xq, xkv, w, b, wo, bo, heads, bias=None):
    """Packed in-proj (q rows, then k, then v); heads are contiguous hd-wide channel chunks."""
    d = xq.shape[-1]
    hd = d // heads
    q = xq @ w[:d].T + b[:d]
    k = xkv @ w[d:2 * d].T + b[d:2 * d]
    v = xkv @ w[2 * d:].T + b[2 * d:]
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.reshape(B, Lq, heads, hd).transpose(1, 2)
    k = k.reshape(B, Lk, heads, hd).transpose(1, 2)
    v = v.reshape(B, Lk, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if bias is not None:
        s = s + bias
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(B, Lq, d) @ wo.T + bo


def decoder_stack(sd, x, mem, tokens, cfg: OracleCfg, prefix="decoder.decoder.layers.", eps=1e-5):
    """x (B,L,d) embedded targets, mem (B,S,d) memory WITH encoder_pos_embed added.
    Mask = causal(-inf above diag) + 1.0 at PAD keys (utils.py:7-12,26-30; Q7)."""
    B, L, d = x.shape
    causal = torch.full((L, L), float("-inf"), dtype=x.dtype, device=x.device).triu(1)
    padbias = (tokens == cfg.pad_idx).to(x.dtype)[:, None, None, :]      # (B,1,1,L) added to keys
    bias = causal[None, None] + padbias
    for i in range(_num_dec_layers(sd, prefix)):
        g = lambda k: sd[f"{prefix}{i}.{k}"]
        x = _ln(x + _mha(x, x, g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias"),
                         g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"),
                         cfg.dec_heads, bias), g("norm1.weight"), g("norm1.bias"), eps)
        x = _ln(x + _mha(x, mem, g("multihead_attn.in_proj_weight"), g("multihead_attn.in_proj_bias"),
                         g("multihead_attn.out_proj.weight"), g("multihead_attn.out_proj.bias"),
                         cfg.dec_heads), g("norm2.weight"), g("norm2.bias"), eps)
        f = torch.relu(x @ g("linear1.weight").T + g("linear1.bias")) @ g("linear2.weight").T + g("linear2.bias")
        x = _ln(x + f, g("norm3.weight"), g("norm3.bias"), eps)
    return x


def decoder_predict(sd, encoder_out, tgt, cfg: OracleCfg):
    """Decoder.predict, model.py:92-127: pad to max_len-1 with PAD, run all positions,
    prepend a constant row of float(bos_idx), drop the last row."""
    B, L = tgt.shape
    pad = torch.full((B, cfg.max_len - 1 - L), cfg.pad_idx, dtype=tgt.dtype, device=tgt.device)
    tokens = torch.cat([tgt, pad], dim=1)
    x = sd["decoder.embedding.weight"][tokens] + sd["decoder.decoder_pos_embed"]
    mem = encoder_out + sd["decoder.encoder_pos_embed"]
    y = decoder_stack(sd, x, mem, tokens, cfg)
    out = y @ sd["decoder.output.weight"].T + sd["decoder.output.bias"]
    bos = torch.full((B, 1, out.shape[-1]), float(cfg.bos_idx), dtype=out.dtype, device=out.device)
    return torch.cat([bos, out[:, :-1]], dim=1)


def interp_pos_embed(pos, length):
    """F.interpolate(mode='linear', align_corners=False) along the sequence axis (model.py:64-68)."""
    n = pos.shape[1]
    if length == n:
        return pos
    # float32 arithmetic exactly as ATen's upsample_linear1d (area_pixel_compute_source_index)
    scale = torch.tensor(n, dtype=torch.float32) / torch.tensor(length, dtype=torch.float32)
    # ATen's compiled kernel contracts scale*(i+0.5)-0.5 into ONE fused multiply-add (single rounding);
    # emulate: the double product of two floats is exact, so rounding the double result once == fmaf
    src = (scale.double() * (torch.arange(length, dtype=torch.float32) + 0.5).double() - 0.5).to(torch.float32)
    src = src.clamp(min=0.0)
    i0 = src.to(torch.int64).clamp(max=n - 1)
    i1 = (i0 + (i0 < n - 1).to(torch.int64))
    w1 = (src - i0.to(torch.float32)).to(pos.dtype)[None, :, None]
    return pos[:, i0] * (1 - w1) + pos[:, i1] * w1


def decoder_forward(sd, encoder_out, tgt, cfg: OracleCfg):
    """Decoder.forward, model.py:58-88: prepend BOS, interpolate pos-embed, return (B,L+1,V)."""
    B = tgt.shape[0]
    tokens = torch.cat([torch.full((B, 1), cfg.bos_idx, dtype=tgt.dtype, device=tgt.device), tgt], dim=1)
    pos = interp_pos_embed(sd["decoder.decoder_pos_embed"], tokens.shape[1])
    x = sd["decoder.embedding.weight"][tokens] + pos
    mem = encoder_out + sd["decoder.encoder_pos_embed"]
    y = decoder_stack(sd, x, mem, tokens, cfg)
    return y @ sd["decoder.output.weight"].T + sd["decoder.output.bias"]


def axial_attention(sd, x, cfg: OracleCfg, axis=-1, prefix="decoder.axial_attention."):
    """AxialAttention.forward, axial_model.py:28-40: un-masked MHA, scale 0.125 regardless of
    head width, to_qkv without bias, to_out with bias, softmax over `axis`."""
    B, n, d = x.shape
    H = cfg.axial_heads
    qkv = x @ sd[prefix + "to_qkv.weight"].T
    q, k, v = [t.reshape(B, n, H, d // H).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
    a = torch.softmax((q @ k.transpose(-1, -2)) * cfg.axial_scale, dim=axis) @ v
    return a.transpose(1, 2).reshape(B, n, d) @ sd[prefix + "to_out.weight"].T + sd[prefix + "to_out.bias"]


def axial_decoder_forward(sd, encoder_out, tgt, cfg: OracleCfg):
    """axial_model.Decoder.forward, axial_model.py:88-118: no BOS prepend; axial attention on the
    raw token embeddings, THEN the positional embedding."""
    pos = interp_pos_embed(sd["decoder.decoder_pos_embed"], tgt.shape[1])
    x = axial_attention(sd, sd["decoder.embedding.weight"][tgt], cfg) + pos
    mem = encoder_out + sd["decoder.encoder_pos_embed"]
    y = decoder_stack(sd, x, mem, tgt, cfg)
    return y @ sd["decoder.output.weight"].T + sd["decoder.output.bias"]


def model_predict(sd, image, tgt, cfg: OracleCfg):
    """EncoderDecoder.predict, model.py:177-181."""
    return decoder_predict(sd, encoder_forward(sd, image, cfg), tgt, cfg)


def model_forward(sd, image, tgt, cfg: OracleCfg):
    """EncoderDecoder.forward in eval mode, model.py:154-175 (patch masking is training-only)."""
    return decoder_forward(sd, encoder_forward(sd, image, cfg), tgt, cfg)


# ----------------------------------------------------------------------------------------
# A.3 incremental decode (what the CUDA path computes): identical by causality
# ----------------------------------------------------------------------------------------
def next_token_logits(sd, encoder_out, prefix_tokens, cfg: OracleCfg):
    """Canonical next-token logits for a prefix of length L := predict(x, prefix)[:, L, :]
    (SURVEY Q5).  Computed without the PAD tail: causal masking makes rows < L independent of it."""
    L = prefix_tokens.shape[1]
    x = sd["decoder.embedding.weight"][prefix_tokens] + sd["decoder.decoder_pos_embed"][:, :L]
    mem = encoder_out + sd["decoder.encoder_pos_embed"]
    y = decoder_stack(sd, x, mem, prefix_tokens, cfg)
    return y[:, -1] @ sd["decoder.output.weight"].T + sd["decoder.output.bias"]


# ----------------------------------------------------------------------------------------
# sampler pieces: inference_p.py:69-90 and transformers' removed top_k_top_p_filtering (Q2)
# ----------------------------------------------------------------------------------------
def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, filter_value=-float("inf"), min_tokens_to_keep=1):
    """Restated from transformers TopKLogitsWarper / TopPLogitsWarper (the function the reference
    imports at inference_p.py:16 no longer exists in transformers 5.x)."""
    logits = logits.clone()
    if top_k > 0:
        k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, filter_value)
    if 0 <= top_p < 1.0:
        srt, idx = torch.sort(logits, descending=False)
        cum = srt.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1 - top_p)
        remove[..., -min_tokens_to_keep:] = False
        logits = logits.masked_fill(remove.scatter(-1, idx, remove), filter_value)
    return logits


def sample_from_uniform(logits, u):
    """Inverse-CDF draw over softmax(filtered logits) with a caller-supplied uniform u in [0,1):
    the deterministic stand-in for torch.multinomial (inference_p.py:74) shared with the CUDA path.
    Picks the first index whose inclusive cumulative probability exceeds u."""
    p = torch.softmax(logits.double(), dim=-1)
    cdf = p.cumsum(-1)
    idx = (cdf <= u[:, None].double() * cdf[:, -1:]).sum(-1)
    return idx.clamp(max=logits.shape[-1] - 1)


def generate(sd, image, cfg: OracleCfg, max_len=50, top_k=0, top_p=1.0, uniforms=None,
             recompute_encoder=False, return_logits=False):
    """generate(), inference_p.py:69-90, with the Q5 row selection (`predict(...)[:, L]`).
    recompute_encoder=True reproduces the reference's per-step encoder pass (Q9) for timing."""
    B = image.shape[0]
    toks = torch.full((B, 1), cfg.bos_idx, dtype=torch.long, device=image.device)
    confs, all_logits = [], []
    enc = None if recompute_encoder else encoder_forward(sd, image, cfg)
    for i in range(max_len):
        e = encoder_forward(sd, image, cfg) if recompute_encoder else enc
        logits = next_token_logits(sd, e, toks, cfg)
        if return_logits:
            all_logits.append(logits)
        logits = top_k_top_p_filtering(logits, top_k=top_k, top_p=top_p)
        if i % 4 == 0:
            confs.append(torch.softmax(logits, dim=-1).max(dim=-1)[0])
        if top_k != 0 or top_p != 1:
            nxt = sample_from_uniform(logits, uniforms[:, i])
        else:
            nxt = torch.softmax(logits, dim=-1).argmax(dim=-1)
        toks = torch.cat([toks, nxt.view(-1, 1)], dim=1)
    if return_logits:
        return toks, confs, torch.stack(all_logits, dim=1)
    return toks, confs


# ----------------------------------------------------------------------------------------
# A.4b token codec, decode side    data_processing.py:317-391 (decode), :547-598 (decode_bboxes)
# ----------------------------------------------------------------------------------------
TOK_EOS, TOK_PAD, TOK_SOC, TOK_EOC, LABEL_LO, LABEL_HI = 301, 302, 303, 304, 258, 267


def _dequant(v, num_bins, extent):
    """data_processing.py:258-262 + :551-553: float32(v) / (num_bins-1) * extent, both steps in float32 (numpy weak scalars)."""
    import numpy as np
    return (np.asarray(v).astype("float32") / (num_bins - 1)) * extent


def decode_bboxes(tokens, num_bins=224, width=224, height=224):
    """data_processing.py:556-598 for an int (B,L) batch -> f32 (B, Nmax, 4), zero-row padded, Nmax >= 1.
    Per sequence: start after the first caption-end token (0 if there is none); at a label token (258..267) read the next four
    tokens as a box, keep it when all are in [0,224] and x1 > x0, y1 > y0, advance by 5 either way; stop at EOS; any other token
    advances by 1.  The loop bound `i < len-4` is the reference's."""
    import numpy as np
    out = []
    for seq in tokens.tolist():
        n = len(seq)
        start = seq.index(TOK_EOC) + 1 if TOK_EOC in seq else 0
        boxes, i = [], start
        while i < n - 4:
            tok = seq[i]
            if LABEL_LO <= tok <= LABEL_HI:
                b = seq[i + 1:i + 5]
                if all(0 <= v <= 224 for v in b) and b[2] > b[0] and b[3] > b[1]:
                    boxes.append(b)
                i += 5
            elif tok == TOK_EOS:
                break
            else:
                i += 1
        if boxes:
            a = np.asarray(boxes)
            f = _dequant(a, num_bins, 1.0).astype("float32")
            f[:, [0, 2]] = f[:, [0, 2]] * width
            f[:, [1, 3]] = f[:, [1, 3]] * height
            out.append(torch.tensor(f).float())
        else:
            out.append(torch.zeros(1, 4))
    return torch.nn.utils.rnn.pad_sequence(out, batch_first=True, padding_value=0)


def decode_sequence(tokens, num_bins=224, width=224, height=224):
    """data_processing.py:317-391 for ONE int sequence -> (labels, boxes f32 (n,4), caption token ids).  PAD tokens are removed
    first, then everything from the first EOS on; a caption needs both 303 and 304; boxes are read in fixed groups of five after
    the caption-end token and kept when the label is 258..267 and all four values are in [0,224] (no ordering test here)."""
    import numpy as np
    seq = [t for t in tokens.tolist() if t != TOK_PAD]
    if TOK_EOS in seq:
        seq = seq[:seq.index(TOK_EOS)]
    labels, boxes, caption = [], [], None          # caption None: no 303/304 pair (the reference returns "" then)
    if TOK_SOC in seq and TOK_EOC in seq:
        soc, eoc = seq.index(TOK_SOC), seq.index(TOK_EOC)
        caption = seq[soc + 1:eoc]
        rest = seq[eoc + 1:]
        for i in range(0, len(rest), 5):
            if i + 4 < len(rest):
                lab, b = rest[i], rest[i + 1:i + 5]
                if LABEL_LO <= lab <= LABEL_HI and all(0 <= v <= 224 for v in b):
                    labels.append(lab); boxes.append(b)
    if boxes:
        a = np.asarray(boxes)
        f = np.empty(a.shape, dtype="float32")
        f[:, [0, 2]] = _dequant(a[:, [0, 2]], num_bins, width)
        f[:, [1, 3]] = _dequant(a[:, [1, 3]], num_bins, height)
        bx = torch.tensor(f)
    else:
        bx = torch.zeros(0, 4)
    return labels, bx, caption


# ----------------------------------------------------------------------------------------
# A.5 box scores
# ----------------------------------------------------------------------------------------
def _pair(b1, b2):
    x0 = torch.maximum(b1[:, None, 0], b2[None, :, 0]); y0 = torch.maximum(b1[:, None, 1], b2[None, :, 1])
    x1 = torch.minimum(b1[:, None, 2], b2[None, :, 2]); y1 = torch.minimum(b1[:, None, 3], b2[None, :, 3])
    inter = (x1 - x0).clamp(min=0) * (y1 - y0).clamp(min=0)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    return inter, a1[:, None] + a2[None, :] - inter


def bbox_iou(b1, b2):
    """iou_calcualtions.py:5-40: inter / (union + 1e-6)."""
    inter, union = _pair(b1, b2)
    return inter / (union + 1e-6)


def calculate_iou(b1, b2):
    """iou_bbox.py:3-43: inter / union, no epsilon (0/0 -> nan)."""
    inter, union = _pair(b1, b2)
    return inter / union


def giou_pairwise(b1, b2):
    """iou_calcualtions.py:220-255."""
    inter, union = _pair(b1, b2)
    ex = torch.maximum(b1[:, None, 2], b2[None, :, 2]) - torch.minimum(b1[:, None, 0], b2[None, :, 0])
    ey = torch.maximum(b1[:, None, 3], b2[None, :, 3]) - torch.minimum(b1[:, None, 1], b2[None, :, 1])
    enc = ex * ey
    return inter / union - (enc - union) / enc


def batch_iou(pred, gt):
    """calculate_batch_iou, iou_calcualtions.py:45-56 -> (B,N,M)."""
    return torch.stack([bbox_iou(pred[i], gt[i]) for i in range(pred.shape[0])])


def batch_max_iou(pred, gt):
    """calculate_batch_max_iou, iou_calcualtions.py:59-75: row-max over GT incl. zero rows (Q13)."""
    return batch_iou(pred, gt).max(dim=2)[0]


def batch_max_iou_torchvision(pred, gt):
    """calculate_batch_max_iou_torchvision, iou_calcualtions.py:78-105: torchvision box_iou
    (no epsilon) then nan_to_num(0)."""
    out = []
    for i in range(pred.shape[0]):
        out.append(torch.nan_to_num(calculate_iou(pred[i], gt[i]), nan=0.0).max(dim=1)[0])
    return torch.stack(out)


def iou_loss(pred, gt, min_penalty=0.5):
    """iou_bbox.py:46-63."""
    iou = calculate_iou(pred, gt)
    iou = torch.where(iou > 0, iou, torch.tensor(min_penalty, dtype=iou.dtype))
    return (1 - iou).mean()


def giou_loss_with_scores(pred, gt, no_detection_penalty=1.0):
    """iou_calcualtions.py:165-208: per image drop all-zero-sum rows, 1 - mean(GIoU);
    penalty = #GT when nothing was predicted; 0 when there is no GT."""
    losses, scores = [], []
    for i in range(pred.shape[0]):
        p = pred[i][pred[i].sum(dim=1) != 0]
        g = gt[i][gt[i].sum(dim=1) != 0]
        if len(p) == 0 and len(g) > 0:
            losses.append(torch.tensor(float(no_detection_penalty * len(g)))); scores.append(torch.zeros(0))
        elif len(p) == 0 or len(g) == 0:
            losses.append(torch.tensor(0.0)); scores.append(torch.zeros(0))
        else:
            s = giou_pairwise(p, g)
            losses.append(1 - s.mean()); scores.append(s)
    return torch.stack([l.to(torch.float32) for l in losses]).mean(), scores


# ----------------------------------------------------------------------------------------
# synthetic inputs: NEU-DET-shaped u8 grayscale -> model tensor
# ----------------------------------------------------------------------------------------
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def synth_gray_u8(B, hw=200, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, hw, hw), generator=g, dtype=torch.uint8)


def _cv2_linear_coeffs(src, dst, zero_outside):
    """OpenCV imgproc/resize.cpp, 8-bit INTER_LINEAR tables: f = (float)((d + 0.5) * scale - 0.5) with scale = 1 / (dst / src) in
    double, s = floor(f), fixed-point coefficients saturate_cast<short>(c * 2048) rounded half-to-even.  In x an index outside the
    row gets (2048, 0) on the clamped pixel (`zero_outside`); in y the coefficients are kept and the two ROWS are clamped."""
    scale = 1.0 / (float(dst) / float(src))
    s_out = np.zeros(dst, np.int64); a = np.zeros((dst, 2), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f)); f = np.float32(f - np.float32(s))
        if zero_outside:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= src - 1:
                f, s = np.float32(0), src - 1
        s_out[d] = s
        a[d, 0] = int(np.clip(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))), -32768, 32767))
        a[d, 1] = int(np.clip(np.rint(np.float32(f * np.float32(2048))), -32768, 32767))
    return s_out, a


def cv2_resize_linear_u8(img, W, H):
    """cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR) for a uint8 (h,w) or (h,w,c) numpy image, restated in integer
    arithmetic -- what A.Resize(size, size) runs on the reference's uint8 images (inference_p.py:148, dataset.py:109-113).
    Pinned bit-exact against cv2 4.13 in oracle/make_golden.py and tests/test_oracle_cpu.py."""
    h, w = img.shape[:2]
    sx, ax = _cv2_linear_coeffs(w, W, True)
    sy, ay = _cv2_linear_coeffs(h, H, False)
    im = img.astype(np.int64)
    x1 = np.minimum(sx + 1, w - 1)
    shx = (1, -1, 1) if img.ndim == 3 else (1, -1)
    rows = im[:, sx] * ax[:, 0].reshape(shx) + im[:, x1] * ax[:, 1].reshape(shx)          # horizontal pass, int32 range
    y0 = np.clip(sy, 0, h - 1); y1 = np.clip(sy + 1, 0, h - 1)
    shy = (-1, 1, 1) if img.ndim == 3 else (-1, 1)
    b0 = ay[:, 0].reshape(shy); b1 = ay[:, 1].reshape(shy)
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2       # vertical pass (VResizeLinear, uchar)
    return np.clip(out, 0, 255).astype(np.uint8)


def albu_normalize(img_u8_hwc):
    """albumentations.Normalize() defaults (functional.normalize, 1.x; third party, absent here -- restated, parity unpinned):
    float32 mean * 255, reciprocal of float32 std * 255, img.astype(float32); img -= mean; img *= denominator."""
    mean = np.array(IMAGENET_MEAN, dtype=np.float32); mean *= 255.0
    std = np.array(IMAGENET_STD, dtype=np.float32); std *= 255.0
    den = np.reciprocal(std, dtype=np.float32)
    img = img_u8_hwc.astype(np.float32)
    img -= mean
    img *= den
    return img


def preprocess_u8(batch_u8, size=224):
    """The reference's VOCDatasetTest transform (inference_p.py:145-158) for a batch: uint8 (B,h,w) gray images (cv2.imread gives
    three equal channels) or (B,h,w,3) BGR images -> f32 (B,3,size,size): [..., ::-1] -> A.Resize -> A.Normalize -> permute(2,0,1)."""
    arr = batch_u8.numpy()
    outs = []
    for im in arr:
        rgb = np.repeat(im[:, :, None], 3, axis=2) if im.ndim == 2 else im[..., ::-1]
        r = cv2_resize_linear_u8(np.ascontiguousarray(rgb), size, size)
        outs.append(torch.from_numpy(albu_normalize(r)).permute(2, 0, 1))
    return torch.stack(outs)


def preprocess_gray(u8, size=224):
    return preprocess_u8(u8, size)


def synthetic_model_inputs(u8, size=224):
    """Generator of the seeded f32 (B,3,size,size) MODEL INPUTS the committed model goldens (case_P_*, case_S_*, case_axial) were
    made with: float bilinear interpolation of the synthetic gray images + ImageNet normalisation.  NOT the reference transform
    (that is preprocess_u8: uint8 cv2 resize) -- just a fixed, smooth-ish input distribution; kept so the goldens stay valid."""
    x = u8.to(torch.float32)[:, None]
    x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False)
    x = x / 255.0
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (x.expand(-1, 3, -1, -1) - mean) / std


def synth_bgr_u8(B, h=260, w=300, seed=99):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, h, w, 3), generator=g, dtype=torch.uint8)


def to_dtype(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------
# detection metrics.  torchmetrics / pycocotools / nltk are absent here and not vendored by the reference
# (train_val_epoch.py:12, utils.py:3): restated from their published algorithms -- PARITY UNPINNED.
# ----------------------------------------------------------------------------------------
def _iou_f64(a, b):
    w = min(a[2], b[2]) - max(a[0], b[0]); h = min(a[3], b[3]) - max(a[1], b[1])
    if w <= 0 or h <= 0:
        return 0.0
    inter = w * h
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter)


def mean_average_precision(preds, targets, iou_thresholds=(0.3,), max_det=100):
    """torchmetrics MeanAveragePrecision(box_format='xyxy', iou_thresholds=[...]).compute()['map'] as the reference uses it
    (train_val_epoch.py:205-231): pycocotools COCOeval evaluateImg (greedy matching per image and class in score order) +
    accumulate (101 recall thresholds, area 'all', maxDets 100), mean over thresholds and classes with ground truth."""
    classes = sorted({int(c) for t in targets for c in t["labels"].tolist()} | {int(c) for p in preds for c in p["labels"].tolist()})
    rec_thrs = np.linspace(0.0, 1.0, 101).tolist()        # COCOeval Params.recThrs: the linspace values, not i / 100
    vals = []
    for thr in iou_thresholds:
        for c in classes:
            dets, npig = [], 0
            for p, t in zip(preds, targets):
                gts = [[float(v) for v in b] for b, l in zip(t["boxes"].tolist(), t["labels"].tolist()) if int(l) == c]
                npig += len(gts)
                d = [([float(v) for v in b], float(s)) for b, s, l in zip(p["boxes"].tolist(), p["scores"].tolist(), p["labels"].tolist()) if int(l) == c]
                d = sorted(d, key=lambda x: -x[1])[:max_det]          # python's sort is stable, like numpy mergesort
                taken = [False] * len(gts)
                for box, score in d:
                    best, m = min(thr, 1 - 1e-10), -1
                    for j, g in enumerate(gts):
                        if taken[j]:
                            continue
                        v = _iou_f64(box, g)
                        if v < best:
                            continue
                        best, m = v, j
                    if m >= 0:
                        taken[m] = True
                    dets.append((score, m >= 0))
            if npig == 0:
                continue
            order = sorted(range(len(dets)), key=lambda i: -dets[i][0])
            tp = fp = 0.0
            rc, pr = [], []
            for i in order:
                if dets[i][1]:
                    tp += 1
                else:
                    fp += 1
                rc.append(tp / npig); pr.append(tp / (fp + tp + np.spacing(1)))
            for i in range(len(pr) - 1, 0, -1):
                if pr[i] > pr[i - 1]:
                    pr[i - 1] = pr[i]
            q = []
            for r in rec_thrs:
                k = next((i for i, x in enumerate(rc) if x >= r), None)
                q.append(pr[k] if k is not None else 0.0)
            vals.append(sum(q) / len(q))
    return sum(vals) / len(vals) if vals else -1.0
