"""TEST INFRASTRUCTURE ONLY -- the seeded parity cases shared by oracle/make_golden.py (which runs the
UNMODIFIED reference on them, in the build container) and tests/ (which run the CUDA path and the
oracle restatement on them, on the GPU box where /root/reference does not exist).

Weights come from the PRODUCT constructors under torch.manual_seed (CPU RNG, deterministic across
hosts for one torch build), optionally with LayerScale gammas redrawn from U(0.5,1.5) -- random-init
DeiT-III has gamma = 1e-6, which hides encoder bugs (SURVEY 7 "hard parts").
"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import mdc_oracle as O  # noqa: E402

VIT = "deit3_medium_patch16_224.fb_in22k_ft_in1k"

# name -> model geometry (SURVEY 8 "Model configurations")
CONFIGS = {
    "P": dict(dim=256, heads=8, layers=6, vocab=305, max_len=100),      # inference_p.py:126-129
    "S": dict(dim=64, heads=2, layers=2, vocab=305, max_len=100),       # inference_trail_after_good_map.py:134-136
    "T": dict(dim=1024, heads=8, layers=8, vocab=332, max_len=100),     # trail_01.py:158-160
}


def product_cfg(max_len=100, pad_idx=302, bos_idx=300):
    import mdcnet_b200 as M
    M.CFG.max_len, M.CFG.pad_idx, M.CFG.bos_idx = max_len, pad_idx, bos_idx
    return M


def build_product_model(config="P", seed=0, gamma_seed=5, axial=False, max_len=None):
    """Product EncoderDecoder on CPU (parameter container only; no forward on CPU)."""
    c = CONFIGS[config]
    M = product_cfg(max_len or c["max_len"])
    torch.manual_seed(seed)
    enc = M.Encoder(model_name=VIT, pretrained=False, out_dim=c["dim"])
    if axial:
        dec = M.axial_model.Decoder(c["vocab"], 196, c["dim"], c["heads"], c["layers"])
        model = M.axial_model.EncoderDecoder(enc, dec)
    else:
        dec = M.Decoder(c["vocab"], 196, c["dim"], c["heads"], c["layers"])
        model = M.EncoderDecoder(enc, dec)
    if gamma_seed is not None:
        g = torch.Generator().manual_seed(gamma_seed)
        with torch.no_grad():
            for k, p in model.named_parameters():
                if k.endswith("gamma"):
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5)
    return model.eval()


def oracle_cfg(config="P", max_len=None):
    c = CONFIGS[config]
    return O.OracleCfg(max_len=max_len or c["max_len"], dec_heads=c["heads"], out_dim=c["dim"])


def state_dict_of(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def images(B, seed=1234):
    return O.synthetic_model_inputs(O.synth_gray_u8(B, seed=seed))


PREFIX = torch.tensor([[300, 5, 302, 17], [300, 260, 100, 302]])      # PAD (302) inside the prefix exercises Q7


def iou_boxes(B=6, N=7, M=5, seed=11):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(B, N, 4, generator=g) * 160
    p[..., 2:] = p[..., :2] + 8 + torch.rand(B, N, 2, generator=g) * 56
    q = torch.rand(B, M, 4, generator=g) * 160
    q[..., 2:] = q[..., :2] + 8 + torch.rand(B, M, 2, generator=g) * 56
    p[0, 4:] = 0; p[3] = 0; q[1, 3:] = 0; q[4] = 0          # pad_sequence-style zero rows, an empty image each side
    p[5, 0] = q[5, 0]                                        # an exact match
    return p, q


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def token_sequences(seed=11):
    """Token batches for the decode-side codec (Tokenizer.decode_bboxes / decode): well-formed sequences in the layout of
    data_processing.py:264-290, damaged ones (missing markers, early EOS, PAD inside, invalid or truncated groups, the literal
    bound 224, x1 <= x0), and pure noise as a random-init model emits it.  int64 (B, 100)."""
    g = torch.Generator().manual_seed(seed)
    L, rows = 100, []

    def ri(lo, hi, n=None):
        return torch.randint(lo, hi, (n,) if n else (), generator=g).tolist()

    def pad(seq):
        seq = seq[:L]
        return seq + [302] * (L - len(seq))

    for k in range(24):                                   # well-formed, 1..6 boxes, 0..14 caption words
        seq = [300, 303] + [270 + w for w in ri(0, 13, int(ri(0, 15)))] + [304]
        for _ in range(int(ri(1, 7))):
            x0, y0 = ri(0, 200), ri(0, 200)
            seq += [258 + ri(0, 10), x0, y0, x0 + ri(1, 24), y0 + ri(1, 24)]
        rows.append(pad(seq + [301]))
    rows.append(pad([300, 258, 10, 20, 30, 40, 301]))                         # no caption markers at all
    rows.append(pad([300, 303, 270, 258, 10, 20, 30, 40, 301]))               # caption start without end
    rows.append(pad([300, 304, 303, 258, 10, 20, 30, 40, 259, 1, 2, 3, 4, 301]))   # end before start
    rows.append(pad([300, 303, 271, 304, 258, 10, 20, 10, 40, 259, 5, 6, 7, 8, 301]))       # x1 == x0 (bboxes drops, decode keeps)
    rows.append(pad([300, 303, 271, 304, 258, 0, 0, 224, 224, 260, 1, 1, 225, 9, 301]))     # literal bound 224 / out of range 225
    rows.append(pad([300, 303, 271, 304, 258, 10, 20, 30, 301, 259, 1, 2, 3, 4]))           # EOS inside a group
    rows.append(pad([300, 303, 271, 304, 302, 258, 10, 20, 30, 40, 302, 259, 1, 2, 3, 4, 301]))   # PADs between groups
    rows.append(pad([300, 303, 271, 304, 7, 8, 258, 10, 20, 30, 40, 301]))                  # stray tokens before the label
    rows.append(pad([300, 303, 271, 304, 299, 10, 20, 30, 40, 258, 1, 2, 3, 4, 301]))       # non-label where a label belongs
    rows.append([300, 303, 271, 304] + [258, 1, 2, 3, 4] * 19 + [258])                      # full length, truncated last group
    rows.append([300, 303] + [270] * 10 + [304] + [259, 5, 6, 50, 60] * 17 + [301, 302])    # EOS near the end
    rows.append(pad([301]))                                                                 # EOS first
    rows.append([302] * L)                                                                   # all PAD
    noise = torch.randint(0, 305, (19, L), generator=g)
    noise[:, 0] = 300
    noise[::2, 7] = 304; noise[::3, 3] = 303
    return torch.cat([torch.tensor(rows, dtype=torch.int64), noise], dim=0)
