"""Where does the bf16 logit error come from?  fp32 path with selected parameters rounded to bf16."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
for gamma, gname in [(None, "case_P_init.pt"), (5, "case_P_gamma.pt")]:
    g = torch.load("tests/golden/" + gname)
    x = cases.images(2).to("cuda")
    n = g["logits"].shape[1]
    def run(select, precision="fp32"):
        m = cases.build_product_model("P", seed=0, gamma_seed=gamma)
        with torch.no_grad():
            for k, p in m.named_parameters():
                if p.dim() > 1 and select(k):
                    p.copy_(p.to(torch.bfloat16).float())
        m = m.to("cuda").set_precision(precision)
        full = m.predict(x, g["tokens"][:, :n].to("cuda"))
        d = (full[:, 1:n + 1].cpu() - g["logits"]).abs()
        return d.max().item(), d.mean().item()
    print(gname)
    print("  fp32 all                       max %.2e mean %.2e" % run(lambda k: False))
    print("  round decoder matrices         max %.2e mean %.2e" % run(lambda k: k.startswith("decoder.decoder") or k.startswith("decoder.output")))
    print("  round decoder.output only      max %.2e mean %.2e" % run(lambda k: k.startswith("decoder.output")))
    print("  round ffn only                 max %.2e mean %.2e" % run(lambda k: "linear" in k))
    print("  round attn proj only           max %.2e mean %.2e" % run(lambda k: "attn" in k and k.startswith("decoder")))
    print("  round encoder matrices         max %.2e mean %.2e" % run(lambda k: k.startswith("encoder")))
    print("  full bf16 path                 max %.2e mean %.2e" % run(lambda k: False, "bf16"))
