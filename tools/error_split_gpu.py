"""Where does the bf16 logit error come from on the GPU path?  Mixes fp32 / bf16 encoder and decoder."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
for gamma, gname in [(None, "case_P_init.pt"), (5, "case_P_gamma.pt")]:
    g = torch.load("tests/golden/" + gname)
    x = cases.images(2).to("cuda")
    n = g["logits"].shape[1]
    toks = g["tokens"][:, :n].to("cuda")
    m = cases.build_product_model("P", seed=0, gamma_seed=gamma).to("cuda")
    def err(full):
        d = (full[:, 1:n + 1].cpu() - g["logits"]).abs()
        return "max %.3e mean %.3e" % (d.max().item(), d.mean().item())
    print(gname)
    for ep in ("fp32", "bf16"):
        m.encoder.set_precision(ep)
        enc = m.encoder(x)
        ref = g["enc_out"]
        e = enc.cpu() if ref.shape[1] == enc.shape[1] else enc.cpu()[:, ::7]
        print(f"  encoder {ep}: enc_out max|d| = {(e - ref).abs().max().item():.3e}  (|enc_out| max {ref.abs().max().item():.2f})")
        for dp in ("fp32", "bf16"):
            m.decoder.set_precision(dp)
            print(f"    encoder {ep} -> decoder {dp}: logits {err(m.decoder.predict(enc, toks))}")
            if dp == "bf16":
                os.environ["MDC_DECODE_BACKEND"] = "generic"
                print(f"    encoder {ep} -> decoder {dp} (generic kernels): logits {err(m.decoder.predict(enc, toks))}")
                os.environ.pop("MDC_DECODE_BACKEND")
    m.set_precision("bf16")
    print(f"  full bf16 model.predict: {err(m.predict(x, toks))}")
