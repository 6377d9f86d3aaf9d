"""CPU-side error budget of the bf16 contract (no GPU needed): the fp32 oracle with selected tensors rounded to
bf16, against the committed reference goldens.  Tells which storage/rounding choices the 2e-2 max-abs budget pays for."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, mdc_oracle as O

bf = lambda t: t.to(torch.bfloat16).float()
FLAGS = dict(sq=False, sk=False, sv=False, so=False, cq=False, co=False, f1=False, f2=False, w_dec=False, w_head=False, w_ckv=False, mem=False, ckv=False, skv=False, act=False, w_enc=False, hid=False, w_fp16=False, own_exact=False, q_bf=False, p_bf=False, act16=False)

def mha(xq, xkv, w, b, wo, bo, heads, bias=None, cross=False):
    d = xq.shape[-1]; hd = d // heads
    rw = (lambda t: t.half().float()) if FLAGS["w_fp16"] else bf
    wq, wk, wv, woo = w[:d], w[d:2*d], w[2*d:], wo
    if FLAGS["w_dec"]: wq, woo = rw(wq), rw(woo)
    if (FLAGS["w_dec"] and not cross) or (FLAGS["w_ckv"] and cross): wk, wv = rw(wk), rw(wv)
    if not cross:
        if FLAGS["sq"]: wq = rw(wq)
        if FLAGS["sk"]: wk = rw(wk)
        if FLAGS["sv"]: wv = rw(wv)
        if FLAGS["so"]: woo = rw(woo)
    else:
        if FLAGS["cq"]: wq = rw(wq)
        if FLAGS["co"]: woo = rw(woo)
    a = bf if FLAGS["act"] else ((lambda t: t.half().float()) if FLAGS["act16"] else (lambda t: t))
    q = a(xq) @ wq.T + b[:d]
    k = a(xkv) @ wk.T + b[d:2*d]; v = a(xkv) @ wv.T + b[2*d:]
    k_ex, v_ex = k, v
    if (cross and FLAGS["ckv"]) or (not cross and FLAGS["skv"]): k, v = bf(k), bf(v)
    B, Lq, _ = q.shape; Lk = k.shape[1]
    q = q.reshape(B, Lq, heads, hd).transpose(1, 2); k = k.reshape(B, Lk, heads, hd).transpose(1, 2); v = v.reshape(B, Lk, heads, hd).transpose(1, 2)
    if FLAGS["q_bf"]: q = bf(q * (1.4426950408889634 / math.sqrt(hd))) / (1.4426950408889634 / math.sqrt(hd))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if bias is not None: s = s + bias
    if FLAGS["own_exact"] and not cross:
        ke = k_ex.reshape(B, Lk, heads, hd).transpose(1, 2); ve = v_ex.reshape(B, Lk, heads, hd).transpose(1, 2)
        sd_ = (q * ke).sum(-1) / math.sqrt(hd)
        idx = torch.arange(Lq)
        s[:, :, idx, idx] = sd_ + (bias[:, :, idx, idx] if bias is not None else 0)
        p = torch.softmax(s, dim=-1)
        pd = p[:, :, idx, idx]
        o = p @ v + pd[..., None] * (ve - v)
    elif FLAGS["p_bf"]:
        m = s.max(-1, keepdim=True).values
        p = torch.exp(s - m)
        o = (bf(p) @ v) / p.sum(-1, keepdim=True)
    else:
        o = torch.softmax(s, dim=-1) @ v
    return a(o.transpose(1, 2).reshape(B, Lq, d)) @ woo.T + bo

def stack(sd, x, mem, tokens, cfg, prefix="decoder.decoder.layers.", eps=1e-5):
    B, L, d = x.shape
    causal = torch.full((L, L), float("-inf")).triu(1)
    bias = causal[None, None] + (tokens == cfg.pad_idx).float()[:, None, None, :]
    rw = (lambda t: t.half().float()) if FLAGS["w_fp16"] else bf
    a = bf if FLAGS["act"] else ((lambda t: t.half().float()) if FLAGS["act16"] else (lambda t: t))
    if FLAGS["mem"]: mem = bf(mem)
    for i in range(O._num_dec_layers(sd, prefix)):
        g = lambda k: sd[f"{prefix}{i}.{k}"]
        x = O._ln(x + mha(x, x, g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias"), g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"), cfg.dec_heads, bias), g("norm1.weight"), g("norm1.bias"), eps)
        x = O._ln(x + mha(x, mem, g("multihead_attn.in_proj_weight"), g("multihead_attn.in_proj_bias"), g("multihead_attn.out_proj.weight"), g("multihead_attn.out_proj.bias"), cfg.dec_heads, cross=True), g("norm2.weight"), g("norm2.bias"), eps)
        w1, w2 = g("linear1.weight"), g("linear2.weight")
        if FLAGS["w_dec"]: w1, w2 = rw(w1), rw(w2)
        if FLAGS["f1"]: w1 = rw(w1)
        if FLAGS["f2"]: w2 = rw(w2)
        h = torch.relu(a(x) @ w1.T + g("linear1.bias"))
        if FLAGS["hid"] or FLAGS["act"]: h = bf(h)
        if FLAGS["act16"]: h = h.half().float()
        x = O._ln(x + h @ w2.T + g("linear2.bias"), g("norm3.weight"), g("norm3.bias"), eps)
    return x

def logits_along(sd, enc_out, toks, cfg, n):
    x = sd["decoder.embedding.weight"][toks[:, :n]] + sd["decoder.decoder_pos_embed"][:, :n]
    mem = enc_out + sd["decoder.encoder_pos_embed"]
    y = stack(sd, x, mem, toks[:, :n], cfg)
    wo = sd["decoder.output.weight"]
    if FLAGS["w_head"]: wo = (wo.half().float() if FLAGS["w_fp16"] else bf(wo))
    a = bf if FLAGS["act"] else ((lambda t: t.half().float()) if FLAGS["act16"] else (lambda t: t))
    return a(y) @ wo.T + sd["decoder.output.bias"]

with torch.no_grad():
    for gamma, gname in [(None, "case_P_init.pt"), (5, "case_P_gamma.pt")]:
        g = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", gname))
        m = cases.build_product_model("P", seed=0, gamma_seed=gamma)
        sd, cfg = cases.state_dict_of(m), cases.oracle_cfg("P")
        x = cases.images(2)
        enc = O.encoder_forward(sd, x, cfg)
        n = g["logits"].shape[1]
        print(gname, "logit rms", g["logits"].std().item(), "max", g["logits"].abs().max().item())
        def run(**kw):
            for k in FLAGS: FLAGS[k] = False
            FLAGS.update(kw)
            d = (logits_along(sd, enc, g["tokens"], cfg, n) - g["logits"]).abs()
            return "max %.2e mean %.2e" % (d.max().item(), d.mean().item())
        print("  none                    ", run())
        print("  w_dec (loop weights)    ", run(w_dec=True))
        print("  w_head                  ", run(w_head=True))
        print("  w_ckv (cross K/V inproj)", run(w_ckv=True))
        print("  mem bf16                ", run(mem=True))
        print("  cross K/V bf16          ", run(ckv=True))
        print("  self K/V bf16           ", run(skv=True))
        print("  activations bf16        ", run(act=True))
        print("  ffn hidden bf16         ", run(hid=True))
        for k in ["sq","sk","sv","so","cq","co","f1","f2"]:
            print("  only %s bf16            " % k, run(**{k: True}))
        base = dict(ckv=True, skv=True, mem=True, w_head=False, w_ckv=True)
        print("  all but head                       ", run(**base, w_dec=True))
        print("  f1,f2,so,co,sv bf16 (q,k exact)     ", run(**base, f1=True, f2=True, so=True, co=True, sv=True))
        print("  f1,f2 bf16 only (+kv)               ", run(**base, f1=True, f2=True))
        print("  f1,f2,sq,sk,sv bf16 (outs exact)    ", run(**base, f1=True, f2=True, sq=True, sk=True, sv=True, cq=True))
        full = dict(w_dec=True, w_head=True, w_ckv=True, ckv=True, skv=True, mem=True)
        print("  FULL (gpu-like)                      ", run(**full))
        print("  FULL, own k/v exact                  ", run(**full, own_exact=True))
        print("  FULL, head exact                     ", run(**{**full, "w_head": False}))
        print("  FULL, head exact + own exact         ", run(**{**full, "w_head": False}, own_exact=True))
        print("  FULL, self kv exact                  ", run(**{**full, "skv": False}))
        print("  FULL, head exact, self kv exact      ", run(**{**full, "skv": False, "w_head": False}))
        new = dict(w_dec=True, w_head=True, w_fp16=True, ckv=True, skv=True, mem=True)   # current GPU policy (w_ckv is bf16 but w_fp16 flips all: approx)
        print("  NEW policy (fp16 loop weights)       ", run(**new))
        print("  NEW + q bf16                         ", run(**new, q_bf=True))
        print("  NEW + p bf16                         ", run(**new, p_bf=True))
        print("  NEW + q bf16 + p bf16                ", run(**new, q_bf=True, p_bf=True))
        print("  NEW + q,p bf16 + activations fp16    ", run(**new, q_bf=True, p_bf=True, act16=True))
        print("  all weights bf16        ", run(w_dec=True, w_head=True, w_ckv=True))
        print("  all weights + kv storage", run(w_dec=True, w_head=True, w_ckv=True, ckv=True, skv=True))
        print("  ... + mem bf16          ", run(w_dec=True, w_head=True, w_ckv=True, ckv=True, skv=True, mem=True))
        print("  ... + act bf16          ", run(w_dec=True, w_head=True, w_ckv=True, ckv=True, skv=True, mem=True, act=True))
        print("  w_dec bf16, head+ckv w fp32, kv storage", run(w_dec=True, ckv=True, skv=True))
        print("  all weights fp16 + kv bf16 + mem bf16 ", run(w_dec=True, w_head=True, w_ckv=True, ckv=True, skv=True, mem=True, w_fp16=True))
