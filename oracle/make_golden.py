"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference files
(/root/reference/model.py, axial_model.py, iou_calcualtions.py, iou_bbox.py, data_processing.py through oracle/shims) on the
seeded cases of oracle/cases.py, and asserts on the way that the restatement oracle/mdc_oracle.py agrees.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
The generate() loop is the reference's (inference_p.py:69-90) with the ONE row-selection fix Q5
(`model.predict(x, prefix)[:, L]`) and the restated top_k_top_p_filtering (Q2).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, mdc_oracle as O, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5


def ref_model(R, product_model, config, axial=False):
    c = cases.CONFIGS[config]
    mod = R["axial_model"] if axial else R["model"]
    with cases.quiet():
        enc = mod.Encoder(model_name=cases.VIT, pretrained=False, out_dim=c["dim"])
        dec = mod.Decoder(c["vocab"], 196, c["dim"], c["heads"], c["layers"])
        m = mod.EncoderDecoder(enc, dec).eval()
    print("  load_state_dict:", m.load_state_dict(product_model.state_dict()))
    return m


def ref_generate(m, x, T):
    """inference_p.py:69-90 with Q5: greedy."""
    toks = torch.full((x.size(0), 1), 300, dtype=torch.long)
    confs, logits = [], []
    with torch.no_grad():
        for i in range(T):
            with cases.quiet():
                preds = m.predict(x, toks)[:, toks.size(1)]
            preds = O.top_k_top_p_filtering(preds, top_k=0, top_p=1.0)
            logits.append(preds.clone())
            if i % 4 == 0:
                confs.append(torch.softmax(preds, dim=-1).sort(axis=-1, descending=True)[0][:, 0])
            nxt = torch.softmax(preds, dim=-1).argmax(dim=-1).view(-1, 1)
            toks = torch.cat([toks, nxt], dim=1)
    return toks, torch.stack(confs, 1), torch.stack(logits, 1)


def close(a, b, what, tol=TOL):
    d = (a - b).abs().max().item()
    print(f"  oracle-vs-reference {what}: max|d| = {d:.3e}")
    assert d <= tol, what


class _Vocab:
    """Stand-in for data_processing.Vocabulary (needs spaCy): id -> word table only, which is all Tokenizer.decode reads."""
    def __init__(self):
        self.itos = {i: f"w{i}" for i in range(270, 299)}


def tokens_case(R):
    """The UNMODIFIED data_processing.Tokenizer (imported through the spacy / albumentations shims) on oracle/cases.token_sequences."""
    import importlib
    print("case tokens")
    dp = importlib.import_module("data_processing")
    tok = dp.Tokenizer(vocab=_Vocab(), num_classes=10, num_bins=224, width=224, height=224, max_len=100)
    seqs = cases.token_sequences()
    with cases.quiet():
        bb = tok.decode_bboxes(seqs)
        dec = [tok.decode(s) for s in seqs]
    ob = O.decode_bboxes(seqs)
    assert bb.shape == ob.shape and torch.equal(bb, ob), "decode_bboxes restatement"
    labels, boxes, caps = [], [], []
    for (lab, bx, cap), s in zip(dec, seqs):
        olab, obx, ocap = O.decode_sequence(s)
        rb = torch.tensor(bx, dtype=torch.float64).reshape(-1, 4)
        assert lab == olab and torch.equal(rb.float(), obx) and torch.equal(rb, rb.float().double()), "decode restatement"
        want_cap = "" if ocap is None else [f"w{t}" if 270 <= t < 299 else "<UNK>" for t in ocap]
        assert cap == want_cap, (cap, want_cap)
        labels.append(lab); boxes.append(rb.float()); caps.append(cap)
    print(f"  {len(seqs)} sequences, {int((bb.abs().sum(-1) != 0).sum())} boxes from decode_bboxes, {sum(len(l) for l in labels)} from decode")
    torch.save({"tokens": seqs, "decode_bboxes": bb, "decode_labels": labels, "decode_boxes": boxes, "decode_captions": caps},
               os.path.join(OUT, "case_tokens.pt"))


def preprocess_case():
    """The reference's VOCDatasetTest transform (inference_p.py:145-158: cv2.imread(...)[..., ::-1] -> A.Resize -> A.Normalize ->
    permute) on synthetic uint8 images with the REAL cv2 (4.13 in this container): `resized_*` are cv2.resize(uint8, INTER_LINEAR)
    outputs -- the integer part of the transform, which the CUDA kernel and the oracle restatement must reproduce bit for bit.
    albumentations is absent (requirements.txt: albumentations>=1.2.1), so A.Normalize is its published arithmetic
    (functional.normalize / normalize_cv2: float32 mean*255 and 1/(std*255)) executed with cv2.subtract / cv2.multiply exactly as
    normalize_cv2 does for 3-channel images, and cross-checked against the numpy form: parity of that step is restated, not pinned."""
    import cv2
    import numpy as np
    print("case preprocess (cv2", cv2.__version__ + ")")
    gray = O.synth_gray_u8(2, seed=4321)                       # (2,200,200) NEU-DET-shaped
    bgr = O.synth_bgr_u8(1, h=150, w=190, seed=99)             # (1,150,190,3) as cv2.imread returns colour images
    small = O.synth_bgr_u8(1, h=64, w=48, seed=5)              # upscaling by a large factor
    big = O.synth_gray_u8(1, hw=300, seed=6)                   # downscaling

    def transform(img_hwc_rgb_u8, size):
        r = cv2.resize(img_hwc_rgb_u8, (size, size), interpolation=cv2.INTER_LINEAR)           # A.Resize on uint8
        assert np.array_equal(r, O.cv2_resize_linear_u8(img_hwc_rgb_u8, size, size)), "cv2.resize restatement"
        mean = np.array(O.IMAGENET_MEAN, dtype=np.float32); mean *= 255.0
        std = np.array(O.IMAGENET_STD, dtype=np.float32); std *= 255.0
        den = np.reciprocal(std, dtype=np.float32)
        f = np.ascontiguousarray(r.astype("float32"))
        cv2.subtract(f, np.array(mean.tolist() + [0], dtype=np.float64), f)                     # normalize_cv2
        cv2.multiply(f, np.array(den.tolist() + [0], dtype=np.float64), f)
        assert np.array_equal(f, O.albu_normalize(r)), "normalize restatement (cv2 form == numpy form)"
        return torch.from_numpy(r.copy()), torch.from_numpy(f).permute(2, 0, 1).contiguous()

    # kept small: the uint8 resize outputs of cv2 (one channel for gray inputs, whose three channels are equal) -- the float32
    # normalisation of exactly these values is asserted above to be O.albu_normalize, which the tests re-apply
    gold = {"gray": gray, "bgr": bgr, "small": small, "big": big}
    for name, batch, size in (("gray", gray, 224), ("bgr", bgr, 224), ("small", small, 224), ("big", big, 224), ("gray320", gray[:1], 320)):
        rs = []
        for im in batch.numpy():
            rgb = np.repeat(im[:, :, None], 3, axis=2) if im.ndim == 2 else np.ascontiguousarray(im[..., ::-1])
            r, o = transform(np.ascontiguousarray(rgb), size)
            assert torch.equal(o, O.preprocess_u8(torch.from_numpy(im)[None], size)[0]), "oracle preprocess_u8"
            rs.append(r[..., 0].contiguous() if im.ndim == 2 else r)
        gold["resized_" + name] = torch.stack(rs)               # uint8: (B,size,size) for gray inputs, (B,size,size,3) RGB for colour
    torch.save(gold, os.path.join(OUT, "case_preprocess.pt"))
    print("  saved case_preprocess.pt:", {k: tuple(v.shape) for k, v in gold.items()})


def main():
    if "--only-preprocess" in sys.argv:          # needs cv2 only, not the reference tree
        os.makedirs(OUT, exist_ok=True)
        return preprocess_case()
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    R = ref_loader.load()
    torch.set_num_threads(8)
    if "--only-tokens" in sys.argv:
        return tokens_case(R)

    # ---- config P, LayerScale gamma ~ U(0.5,1.5) ------------------------------------------------
    print("case P (gamma U(0.5,1.5))")
    pm = cases.build_product_model("P", seed=0, gamma_seed=5)
    rm = ref_model(R, pm, "P")
    sd, cfg = cases.state_dict_of(pm), cases.oracle_cfg("P")
    x = cases.images(2)
    with torch.no_grad(), cases.quiet():
        enc_out = rm.encoder(x)
        pred = rm.predict(x, cases.PREFIX)
        fwd = rm(x, cases.PREFIX[:, 1:])
    close(enc_out, O.encoder_forward(sd, x, cfg), "encoder_out")
    close(pred, O.model_predict(sd, x, cases.PREFIX, cfg), "predict")
    close(fwd, O.model_forward(sd, x, cases.PREFIX[:, 1:], cfg), "forward")
    toks, confs, logits = ref_generate(rm, x, 24)
    otoks, oconfs, ologits = O.generate(sd, x, cfg, max_len=24, return_logits=True)
    assert torch.equal(toks, otoks), "greedy tokens"
    close(logits, ologits, "generate logits")
    close(confs, torch.stack(oconfs, 1), "confs")
    top2 = logits.topk(2, dim=-1)[0]
    print("  min top1-top2 margin over steps:", (top2[..., 0] - top2[..., 1]).min().item())
    torch.save({"enc_out": enc_out, "predict": pred, "forward": fwd, "tokens": toks, "confs": confs,
                "logits": logits, "prefix": cases.PREFIX}, os.path.join(OUT, "case_P_gamma.pt"))

    # ---- config P, as constructed (gamma 1e-6) ---------------------------------------------------
    print("case P (as constructed)")
    pm0 = cases.build_product_model("P", seed=0, gamma_seed=None)
    rm0 = ref_model(R, pm0, "P")
    with torch.no_grad(), cases.quiet():
        enc0 = rm0.encoder(x)
    toks0, confs0, logits0 = ref_generate(rm0, x, 12)
    torch.save({"enc_out": enc0[:, ::7].clone(), "tokens": toks0, "confs": confs0, "logits": logits0},
               os.path.join(OUT, "case_P_init.pt"))

    # ---- config S, full-length greedy decode (T = max_len-1 = 99) --------------------------------
    print("case S (T=98: the longest decode predict(x, prefix)[:, L] can express -- row 99 does not exist)")
    ps = cases.build_product_model("S", seed=1, gamma_seed=6)
    rs = ref_model(R, ps, "S")
    xs = cases.images(3, seed=77)
    toks_s, confs_s, logits_s = ref_generate(rs, xs, 98)
    o_toks, o_confs, o_logits = O.generate(cases.state_dict_of(ps), xs, cases.oracle_cfg("S"), max_len=98, return_logits=True)
    assert torch.equal(toks_s, o_toks)
    close(logits_s, o_logits, "S logits")
    torch.save({"tokens": toks_s, "confs": confs_s, "logits": logits_s[:, ::3].clone()}, os.path.join(OUT, "case_S_T98.pt"))

    # ---- axial variant ----------------------------------------------------------------------------
    print("case axial")
    pa = cases.build_product_model("P", seed=2, gamma_seed=7, axial=True)
    ra = ref_model(R, pa, "P", axial=True)
    sda = cases.state_dict_of(pa)
    g = torch.Generator().manual_seed(3)
    tgt12 = torch.randint(0, 305, (2, 12), generator=g); tgt12[:, 0] = 300; tgt12[1, 7] = 302
    tgt99 = torch.randint(0, 305, (2, 99), generator=g); tgt99[:, 0] = 300
    xa = torch.randn(2, 17, 256, generator=g)
    with torch.no_grad(), cases.quiet():
        f12 = ra(x, tgt12); f99 = ra(x, tgt99)
        ax1 = ra.decoder.axial_attention(xa); ax2 = ra.decoder.axial_attention(xa, axis=-2)
    enc_a = O.encoder_forward(sda, x, cfg)
    close(f12, O.axial_decoder_forward(sda, enc_a, tgt12, cfg), "axial forward L=12")
    close(f99, O.axial_decoder_forward(sda, enc_a, tgt99, cfg), "axial forward L=99")
    close(ax1, O.axial_attention(sda, xa, cfg), "AxialAttention axis=-1")
    close(ax2, O.axial_attention(sda, xa, cfg, axis=-2), "AxialAttention axis=-2")
    torch.save({"tgt12": tgt12, "tgt99": tgt99, "f12": f12, "f99": f99[:, ::9].clone(), "xa": xa, "ax1": ax1, "ax2": ax2},
               os.path.join(OUT, "case_axial.pt"))

    # ---- IoU: the reference's own functions, unmodified ------------------------------------------
    print("case iou")
    I, IB = R["iou_calcualtions"], R["iou_bbox"]
    p, q = cases.iou_boxes()
    gl, gs = I.giou_loss_with_scores(p, q)
    gold = {
        "pred": p, "gt": q,
        "batch_iou": torch.stack(I.calculate_batch_iou(p, q)),
        "max_iou": torch.tensor(I.calculate_batch_max_iou(p, q)),
        "max_iou_tv": torch.tensor(I.calculate_batch_max_iou_torchvision(p, q)),
        "giou_2": I.giou_pairwise(p[2], q[2]),
        "giou_loss": gl, "giou_scores": gs,
        "calc_iou_2": IB.calculate_iou(p[2], q[2]), "calc_iou_zero": IB.calculate_iou(p[3], q[4]),
        "iou_loss_2": IB.iou_loss(p[2], q[2]),
        "kat_iou": I.bbox_iou(torch.tensor([[0., 0, 10, 10]]), torch.tensor([[5., 5, 15, 15]])),
        "kat_giou": I.giou_pairwise(torch.tensor([[0., 0, 10, 10]]), torch.tensor([[5., 5, 15, 15]])),
    }
    assert torch.equal(gold["batch_iou"], O.batch_iou(p, q))
    assert torch.equal(gold["max_iou"], O.batch_max_iou(p, q).flatten())
    assert torch.equal(gold["max_iou_tv"], O.batch_max_iou_torchvision(p, q).flatten())
    assert torch.equal(gold["giou_2"], O.giou_pairwise(p[2], q[2]))
    assert abs(gl.item() - O.giou_loss_with_scores(p, q)[0].item()) < 1e-6
    torch.save(gold, os.path.join(OUT, "case_iou.pt"))
    tokens_case(R)
    preprocess_case()
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
