"""GPU: the whole hot path behind the reference's API against (i) golden vectors produced by the
UNMODIFIED reference and (ii) the oracle restatement on the same seeded weights/inputs.

Contract (BASELINE.json north_star): fp32 path -> greedy token ids identical; bf16 path -> logits
max-abs <= 2e-2 and cosine >= 0.999; IoU within 1e-6."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import mdcnet_b200 as M  # noqa: E402
from oracle import cases, mdc_oracle as O  # noqa: E402
from tests import gpu_util as G  # noqa: E402

DEV = "cuda"
FP32_LOGIT_TOL = 2e-4     # fp32 CUDA vs fp32 CPU reference: summation-order noise only
BF16_MAXABS, BF16_COS = 2e-2, 0.999


@pytest.fixture(scope="module")
def model_p():
    return cases.build_product_model("P", seed=0, gamma_seed=5).to(DEV)


@pytest.fixture(scope="module")
def x2():
    return cases.images(2).to(DEV)


def test_fp32_encoder_predict_forward_vs_reference_golden(model_p, x2, golden):
    g = golden("case_P_gamma.pt")
    model_p.set_precision("fp32")
    enc = model_p.encoder(x2)
    assert (enc.cpu() - g["enc_out"]).abs().max().item() < 1e-4
    pred = model_p.predict(x2, g["prefix"].to(DEV))
    assert pred.shape == (2, 99, 305) and torch.all(pred[:, 0] == 300.0)
    assert (pred.cpu() - g["predict"]).abs().max().item() < FP32_LOGIT_TOL
    fwd = model_p(x2, g["prefix"][:, 1:].to(DEV))
    assert fwd.shape == (2, 4, 305)
    assert (fwd.cpu() - g["forward"]).abs().max().item() < FP32_LOGIT_TOL
    # Decoder-level entry points with a caller-provided encoder_out
    pred2 = model_p.decoder.predict(g["enc_out"].to(DEV), g["prefix"].to(DEV))
    assert (pred2.cpu() - g["predict"]).abs().max().item() < FP32_LOGIT_TOL


def test_fp32_greedy_tokens_exact_vs_reference_golden(model_p, x2, golden):
    g = golden("case_P_gamma.pt")
    model_p.set_precision("fp32")
    toks, confs, logits = model_p.generate_tokens(x2, 24, return_logits=True)
    err = (logits.cpu() - g["logits"]).abs().max().item()
    top2 = g["logits"].topk(2, dim=-1)[0]
    margin = (top2[..., 0] - top2[..., 1]).min().item()
    print(f"fp32 path: max|dlogit| = {err:.2e}; reference min top1-top2 margin = {margin:.2e}")
    assert err < FP32_LOGIT_TOL and err < margin
    assert torch.equal(toks.cpu().long(), g["tokens"])
    assert (confs.cpu() - g["confs"]).abs().max().item() < 1e-5
    # public API: generate() -> CPU LongTensor + list of conf tensors (inference_p.py:90)
    bp, cf = M.generate(model_p, x2, M.Tokenizer(), max_len=24)
    assert bp.dtype == torch.long and bp.device.type == "cpu" and torch.equal(bp, g["tokens"])
    assert len(cf) == 6 and all(c.shape == (2,) for c in cf)


def test_fp32_as_constructed_weights_vs_reference_golden(x2, golden):
    g = golden("case_P_init.pt")
    m = cases.build_product_model("P", seed=0, gamma_seed=None).to(DEV).set_precision("fp32")
    assert (m.encoder(x2).cpu()[:, ::7] - g["enc_out"]).abs().max().item() < 1e-4
    toks, confs, logits = m.generate_tokens(x2, 12, return_logits=True)
    assert (logits.cpu() - g["logits"]).abs().max().item() < FP32_LOGIT_TOL
    assert torch.equal(toks.cpu().long(), g["tokens"])


def test_fp32_full_length_decode_config_S(golden):
    """T = 98 = the longest decode the reference's predict()[:, L] can express; pages cross 6 boundaries."""
    g = golden("case_S_T98.pt")
    m = cases.build_product_model("S", seed=1, gamma_seed=6).to(DEV).set_precision("fp32")
    x = cases.images(3, seed=77).to(DEV)
    toks, confs, logits = m.generate_tokens(x, 98, return_logits=True)
    assert (logits.cpu()[:, ::3] - g["logits"]).abs().max().item() < FP32_LOGIT_TOL
    assert torch.equal(toks.cpu().long(), g["tokens"])
    assert (confs.cpu() - g["confs"]).abs().max().item() < 1e-5
    # superset: one more step than the reference can index (position 98 exists in the pos table)
    toks99, _ = m.generate_tokens(x, 99)
    assert torch.equal(toks99[:, :99].cpu().long(), g["tokens"])
    with pytest.raises(RuntimeError):
        m.generate_tokens(x, 100)                          # Q6: no positional row left


def test_bf16_logits_within_contract(model_p, x2, golden):
    g = golden("case_P_gamma.pt")
    model_p.set_precision("bf16")
    enc = model_p.encoder(x2)
    e_err = (enc.cpu() - g["enc_out"]).abs().max().item()
    pred = model_p.predict(x2, g["prefix"].to(DEV))
    err = (pred.cpu()[:, 1:] - g["predict"][:, 1:]).abs().max().item()
    c = G.cos(pred.cpu()[:, 1:], g["predict"][:, 1:])
    print(f"bf16 path: encoder max|d| = {e_err:.3e}; logits max|d| = {err:.3e}, cosine = {c:.6f}")
    assert err <= BF16_MAXABS and c >= BF16_COS
    # same contract on the generic (unfused) decode kernels
    with M.decode_options(per_op_kernels=True):
        pred_g = model_p.predict(x2, g["prefix"].to(DEV))
    assert (pred_g.cpu()[:, 1:] - g["predict"][:, 1:]).abs().max().item() <= BF16_MAXABS
    # teacher-forced per-step logits along the reference's own greedy trajectory
    toks = g["tokens"].to(DEV)
    full = model_p.predict(x2, toks[:, :24])
    step_logits = full[:, 1:25].cpu()
    err2 = (step_logits - g["logits"]).abs().max().item()
    assert err2 <= BF16_MAXABS and G.cos(step_logits, g["logits"]) >= BF16_COS


def test_prefill_forward_and_predict_bf16_vs_reference_golden(model_p, x2, golden):
    """forward() (BOS prepend + interpolated positional table, model.py:58-88) and predict() through the
    all-positions prefill pass against the unmodified reference's outputs, PAD-in-prefix (Q7) included; and its cost at B = 64."""
    g = golden("case_P_gamma.pt")
    model_p.set_precision("bf16")
    pred = model_p.predict(x2, g["prefix"].to(DEV))
    fwd = model_p(x2, g["prefix"][:, 1:].to(DEV))
    assert pred.shape == (2, 99, 305) and torch.all(pred[:, 0] == 300.0) and fwd.shape == (2, 4, 305)
    ep = (pred.cpu()[:, 1:] - g["predict"][:, 1:]).abs().max().item(); ef = (fwd.cpu() - g["forward"]).abs().max().item()
    print(f"prefill bf16: predict max|d| = {ep:.3e}, forward max|d| = {ef:.3e}")
    assert ep <= BF16_MAXABS and ef <= BF16_MAXABS and G.cos(pred.cpu()[:, 1:], g["predict"][:, 1:]) >= BF16_COS
    with M.decode_options(prefill=False):
        fwd_s = model_p(x2, g["prefix"][:, 1:].to(DEV))
    assert (fwd_s - fwd).abs().max().item() < 1.5e-2
    x = cases.images(64, seed=3).to(DEV)
    toks = torch.randint(0, 300, (64, 99), device=DEV); toks[:, 0] = 300
    eng = model_p._engine(torch.device(DEV))
    _, memory = eng.encode(x, want_enc_out=False, want_memory=True)
    for opts in ({"prefill": True}, {"prefill": False}):
        with M.decode_options(**opts):
            model_p.decoder._predict_with(eng, memory, toks); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); model_p.decoder._predict_with(eng, memory, toks); b.record(); torch.cuda.synchronize()
        print(f"predict() decoder part at B = 64, 99 positions, {opts}: {a.elapsed_time(b):.3f} ms (incl. cross-K/V build)")


def test_bf16_contract_at_the_bench_operating_point():
    """The kernel instantiation the headline runs -- 16 images per cluster, full-length decode -- against the CPU oracle directly:
    B = 16 images, teacher-forced along the ORACLE's own greedy trajectory over the whole positional table (predict() rows 1..98 =
    decode steps 0..97), so a near-tie token flip cannot hide a logit error.  Contract: max-abs <= 2e-2, cosine >= 0.999."""
    m = cases.build_product_model("P", seed=0, gamma_seed=5)
    sd, cfg = cases.state_dict_of(m), cases.oracle_cfg("P")
    x = cases.images(16, seed=404)
    want_toks, _, want_logits = O.generate(sd, x, cfg, max_len=98, return_logits=True)       # (16, 99) tokens, (16, 98, V) logits
    m = m.to(DEV).set_precision("bf16")
    for opts in ({"images_per_cluster": 16, "prefill": False}, {"images_per_cluster": 8, "ctas_per_sm": 2, "prefill": False}, {"prefill": True}):
        with M.decode_options(**opts):          # prefill = False: the autoregressive kernel, teacher-forced; True: the all-positions pass
            full = m.predict(x.to(DEV), want_toks[:, :98].to(DEV))
        got = full[:, 1:99].cpu()
        err, c = (got - want_logits).abs().max().item(), G.cos(got, want_logits)
        print(f"bf16 {opts} vs oracle, B=16, 98 positions: logits max|d| = {err:.3e}, cosine = {c:.6f}")
        assert err <= BF16_MAXABS and c >= BF16_COS, opts


def test_bf16_as_constructed_weights_within_contract(x2, golden):
    """The contract is stated for random-init weights (LayerScale 1e-6); the gamma~U(0.5,1.5) set above is a stress case.
    Error budget (tools/error_budget_cpu.py, tools/error_split_gpu.py): with bf16 decode-loop weights this case sat AT the
    limit (2.03e-2: 1.4e-2 from weight rounding alone); with fp16 decode-loop weights it is well inside."""
    g = golden("case_P_init.pt")
    m = cases.build_product_model("P", seed=0, gamma_seed=None).to(DEV).set_precision("bf16")
    full = m.predict(x2, g["tokens"][:, :12].to(DEV))
    step_logits = full[:, 1:13].cpu()
    err = (step_logits - g["logits"]).abs().max().item()
    print(f"bf16 path, as-constructed weights: logits max|d| = {err:.3e}")
    assert err <= 0.75 * BF16_MAXABS and G.cos(step_logits, g["logits"]) >= 0.9999


def test_axial_variant_vs_reference_golden(x2, golden):
    g = golden("case_axial.pt")
    m = cases.build_product_model("P", seed=2, gamma_seed=7, axial=True).to(DEV).set_precision("fp32")
    ax1 = m.decoder.axial_attention(g["xa"].to(DEV))
    ax2 = m.decoder.axial_attention(g["xa"].to(DEV), axis=-2)
    assert (ax1.cpu() - g["ax1"]).abs().max().item() < 2e-5
    assert (ax2.cpu() - g["ax2"]).abs().max().item() < 2e-5
    f12 = m(x2, g["tgt12"].to(DEV))
    assert f12.shape == (2, 12, 305) and (f12.cpu() - g["f12"]).abs().max().item() < FP32_LOGIT_TOL
    f99 = m(x2, g["tgt99"].to(DEV))
    assert (f99.cpu()[:, ::9] - g["f99"]).abs().max().item() < FP32_LOGIT_TOL
    m.set_precision("bf16")
    f12b = m(x2, g["tgt12"].to(DEV))
    assert (f12b.cpu() - g["f12"]).abs().max().item() <= 3e-2 and G.cos(f12b.cpu(), g["f12"]) >= BF16_COS   # stress weights; cosine is the robust check


def test_topk_sampling_matches_oracle_draws(model_p, x2):
    """top-k=5 (train_val_epoch.py:81) with shared uniforms: token-exact vs the oracle's inverse-CDF draw (fp32 path)."""
    model_p.set_precision("fp32")
    u = torch.rand(2, 16, generator=torch.Generator().manual_seed(4))
    toks, confs = model_p.generate_tokens(x2, 16, top_k=5, uniforms=u.to(DEV))
    sd = cases.state_dict_of(model_p.cpu()); model_p.to(DEV)
    otoks, oconfs = O.generate(sd, x2.cpu(), cases.oracle_cfg("P"), max_len=16, top_k=5, uniforms=u)
    assert torch.equal(toks.cpu().long(), otoks)
    assert (confs.cpu() - torch.stack(oconfs, 1)).abs().max().item() < 1e-5


def test_batch_invariance_at_bench_size(model_p):
    """Config 2 size (B=64, T=99): every image's tokens must equal what it gets alone / in another slot
    (no cross-image leakage through the paged KV pool, cross-K/V or workspace)."""
    model_p.set_precision("bf16")
    x = cases.images(64, seed=5).to(DEV)
    toks, confs = model_p.generate_tokens(x, 99)
    assert toks.shape == (64, 100) and torch.all(toks[:, 0] == 300)
    assert torch.all((toks >= 0) & (toks < 305))
    sub = torch.tensor([63, 0, 17])
    toks_s, confs_s = model_p.generate_tokens(x[sub], 99)
    assert torch.equal(toks_s, toks[sub])
    assert torch.equal(confs_s, confs[sub])


def test_load_state_dict_invalidates_prepared_weights(x2):
    m = cases.build_product_model("S", seed=1, gamma_seed=6).to(DEV).set_precision("fp32")
    a = m.predict(x2, torch.tensor([[300], [300]], device=DEV))
    other = cases.build_product_model("S", seed=9, gamma_seed=6)
    m.load_state_dict(other.state_dict())
    b = m.predict(x2, torch.tensor([[300], [300]], device=DEV))
    assert (a - b).abs().max().item() > 1e-3
    want = O.model_predict(cases.state_dict_of(other), x2.cpu(), torch.tensor([[300], [300]]), cases.oracle_cfg("S"))
    assert (b.cpu() - want).abs().max().item() < 2e-4


def test_input_size_mismatch_raises_like_timm(model_p):
    with pytest.raises(AssertionError):
        model_p.encoder(torch.zeros(1, 3, 200, 200, device=DEV))


@pytest.mark.parametrize("B,T", [(2, 24), (5, 40), (37, 30), (64, 99)])
def test_cluster_decode_kernel_matches_generic_kernels(model_p, golden, B, T):
    """The persistent cluster-cooperative decode kernel (decode_cluster.cu) against the unfused kernels (decode.cu) on the
    same weights and KV precision.  Teacher-forced along the generic kernels' own greedy trajectory, so that a near-tie
    token flip cannot hide (or fake) a logit difference; then the free-running trajectories are compared."""
    import os
    model_p.set_precision("bf16")
    x = cases.images(B, seed=21).to(DEV)
    with M.decode_options(per_op_kernels=True):
        tg, _ = model_p.generate_tokens(x, T, use_graph=False)
        lg = model_p.predict(x, tg[:, :T].long())[:, 1:T + 1]
    with M.decode_options(prefill=False):
        lc = model_p.predict(x, tg[:, :T].long())[:, 1:T + 1]
    tc, _ = model_p.generate_tokens(x, T, use_graph=False)
    lp = model_p.predict(x, tg[:, :T].long())[:, 1:T + 1]            # the all-positions prefill pass on the same tokens
    perr = (lp - lc).abs().max().item()
    print(f"prefill vs autoregressive kernel: B={B} T={T} max|dlogit| = {perr:.2e}")
    assert perr < 1.5e-2
    err = (lg - lc).abs().max().item()
    mean = (lg - lc).abs().mean().item()
    agree = (tg == tc).float().mean().item()
    print(f"cluster vs generic: B={B} T={T} teacher-forced max|dlogit| = {err:.2e} mean = {mean:.2e}; free-running token agreement {agree:.3f}")
    # the generic kernels keep fp32 activations / queries / probabilities, the fused kernel rounds projection operands to fp16 and
    # attention queries / probabilities to bf16: both sit ~1e-2 (worst case) from the fp32 reference
    assert err < 1.5e-2 and mean < 1.5e-3
    assert agree > 0.5, agree


def test_cluster_decode_with_runaway_attention_scores_stays_finite_and_consistent():
    """The fused kernel's attention keeps the maximum of an image's FIRST key tile as the softmax reference and raises it (rescaling the
    running state) only when a later score exceeds it by 64 log2-units.  With ordinary weights that path never runs, so force it: the
    attention in-projections of two layers scaled by 16 (scores x 256: one-hot-like softmax, score ranges of hundreds of nats).
    The 8- and 16-image instantiations must stay finite and bitwise equal; against the per-operation kernels (true maximum, fp32
    queries) only a loose bound holds -- a bf16 rounding of a score of 300 nats moves a softmax weight by e^0.6."""
    m = cases.build_product_model("P", seed=3, gamma_seed=5)
    with torch.no_grad():
        for l in (0, 3):
            layer = m.decoder.decoder.layers[l]
            layer.self_attn.in_proj_weight[:512] *= 16.0          # q and k rows
            layer.multihead_attn.in_proj_weight[:512] *= 16.0
    m = m.to(DEV).set_precision("bf16")
    B, T = 16, 40
    x = cases.images(B, seed=33).to(DEV)
    with M.decode_options(per_op_kernels=True):
        tg, _ = m.generate_tokens(x, T, use_graph=False)
        lg = m.predict(x, tg[:, :T].long())[:, 1:T + 1]
    outs = []
    for ipc in (8, 16):
        with M.decode_options(prefill=False, images_per_cluster=ipc):
            outs.append(m.predict(x, tg[:, :T].long())[:, 1:T + 1])
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])
    err, mean = (lg - outs[0]).abs().max().item(), (lg - outs[0]).abs().mean().item()
    print(f"runaway scores: fused vs per-operation kernels max|dlogit| = {err:.2e}, mean = {mean:.2e}")
    assert mean < 3e-2 and G.cos(outs[0], lg) > 0.999, (err, mean)


def test_softmax_rescale_path_against_a_running_maximum_in_the_developer_build():
    """tools/ref_margin_check.py: the developer build of the library reads the reference-maximum margin from the environment; margin 0
    turns the same code into a running maximum (the rescale path on every new maximum).  Both are exact softmax evaluations."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dev = os.path.join(root, "mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200", "libmdc_b200_dev.so")
    if not os.path.exists(dev):
        pytest.skip("developer build (build.py --devtools) not present")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "ref_margin_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-400:], r.stderr[-400:])
    assert r.returncode == 0 and "ok" in r.stdout


@pytest.mark.parametrize("B,T", [(13, 30), (37, 40), (64, 99)])
def test_cluster_decode_16_images_per_cluster_is_bitwise_the_8_image_kernel(model_p, B, T):
    """The two-column-block instantiation of the fused kernel (up to 16 images per cluster pass: what the batch pipeline asks
    for) against the one-block instantiation (up to 8): every image's arithmetic is the same sequence of operations, so logits,
    tokens and confidences must be bitwise equal -- teacher-forced, free-running greedy and top-k sampling."""
    import os
    model_p.set_precision("bf16")
    x = cases.images(B, seed=77).to(DEV)
    u = torch.rand(B, T, generator=torch.Generator().manual_seed(5)).to(DEV)
    def run():
        t, c = model_p.generate_tokens(x, T, use_graph=False)
        ts, cs = model_p.generate_tokens(x, T, top_k=5, uniforms=u, use_graph=False)
        return t, c, ts, cs, model_p.predict(x, t[:, :T].long())
    with M.decode_options(prefill=False):
        want = run()
    for opts in ({"images_per_cluster": 16, "prefill": False}, {"images_per_cluster": 8, "ctas_per_sm": 2, "prefill": False}):
        with M.decode_options(**opts):
            got = run()
        for a, b in zip(got, want):
            assert torch.equal(a, b), opts


def test_greedy_select_breaks_ties_like_torch_argmax():
    """Three vocabulary rows (in three different CTAs of the cluster) carry identical weights and a large bias, so their logits
    tie exactly at every step: the greedy select must return the lowest index (torch.argmax, inference_p.py:77) and the max-prob
    must equal the reference formula on the kernel's own logits -- fused kernel (one image per cluster slot and sixteen) and the
    generic select kernel alike."""
    import os
    m = cases.build_product_model("P", seed=0, gamma_seed=5)
    with torch.no_grad():
        w, b = m.decoder.output.weight, m.decoder.output.bias
        for j in (150, 290):
            w[j].copy_(w[7]); b[j].copy_(b[7])
        b[[7, 150, 290]] += 8.0
    m = m.to(DEV).set_precision("bf16")
    x = cases.images(19, seed=12).to(DEV)
    for env in ({}, {"images_per_cluster": 16}, {"ctas_per_sm": 2}, {"per_op_kernels": True}):
        with M.decode_options(**env):
            toks, confs, logits = m.generate_tokens(x, 12, return_logits=True, use_graph=False)
        assert torch.equal(logits[..., 7], logits[..., 150]) and torch.equal(logits[..., 7], logits[..., 290])
        assert (toks[:, 1:] == 7).all(), (env, toks[0])
        want = torch.softmax(logits.float(), -1).max(-1)[0][:, ::4]
        assert (confs[:, :want.shape[1]] - want).abs().max().item() < 1e-6, env


def test_cluster_decode_is_run_to_run_deterministic(model_p):
    """Race detector for the fused kernel: teacher-forced (no select / token exchange between steps -- the path where a missing
    barrier between the head MMAs and the next step's operand write once showed up) and free-running, 6 runs each, bitwise."""
    model_p.set_precision("bf16")
    x = cases.images(64, seed=33).to(DEV)
    toks, confs = model_p.generate_tokens(x, 99, use_graph=False)
    with M.decode_options(prefill=False):
        ref = model_p.predict(x, toks[:, :99].long())
    pre = model_p.predict(x, toks[:, :99].long())                 # prefill pass: deterministic too
    for _ in range(5):
        with M.decode_options(prefill=False):
            assert torch.equal(model_p.predict(x, toks[:, :99].long()), ref)
        assert torch.equal(model_p.predict(x, toks[:, :99].long()), pre)
        t2, c2 = model_p.generate_tokens(x, 99, use_graph=False)
        assert torch.equal(t2, toks) and torch.equal(c2, confs)


def test_cluster_decode_topk_and_graph_replay(model_p):
    model_p.set_precision("bf16")
    x = cases.images(8, seed=3).to(DEV)
    u = torch.rand(8, 20, generator=torch.Generator().manual_seed(9)).to(DEV)
    a, _ = model_p.generate_tokens(x, 20, top_k=5, uniforms=u)                  # captured graph
    b, _ = model_p.generate_tokens(x, 20, top_k=5, uniforms=u)                  # replay
    c, _ = model_p.generate_tokens(x, 20, top_k=5, uniforms=u, use_graph=False)  # eager
    assert torch.equal(a, b) and torch.equal(a, c)
    g, _ = model_p.generate_tokens(x, 20)
    assert not torch.equal(a, g)


def test_batch_pipeline_equals_serial_generate(model_p):
    """GenerationPipeline / generate_stream (encoder of batch i+1 overlapping the decode loop of batch i on a second stream)
    must return exactly what the serial generate() returns, batch by batch, including a ragged last batch and sampling."""
    model_p.set_precision("bf16")
    tok = M.Tokenizer()
    xs = [cases.images(16, seed=200 + i) for i in range(5)] + [cases.images(5, seed=300)]
    want = [M.generate(model_p, x.to(DEV), tok, max_len=40) for x in xs]
    got = list(M.generate_stream(model_p, (x.pin_memory() for x in xs), tok, max_len=40))
    assert len(got) == len(want)
    for (gt, gc), (wt, wc) in zip(got, want):
        assert torch.equal(gt, wt) and len(gc) == len(wc) and all(torch.equal(a, b) for a, b in zip(gc, wc))
    # second call reuses the cached plans; device inputs; explicit pipeline with top-k sampling and shared uniforms
    got2 = list(M.generate_stream(model_p, (x.to(DEV) for x in xs), tok, max_len=40))
    assert all(torch.equal(a[0], b[0]) for a, b in zip(got2, want))
    pipe = M.GenerationPipeline(model_p, 16, 24, top_k=5, depth=3, decode_streams=2)
    us = [torch.rand(16, 24, generator=torch.Generator().manual_seed(i)).to(DEV) for i in range(4)]
    tickets = [pipe.submit(xs[i].to(DEV), uniforms=us[i]) for i in range(4)]
    pipe.join()
    for i, t in enumerate(tickets):
        ref, _ = model_p.generate_tokens(xs[i].to(DEV), 24, top_k=5, uniforms=us[i])
        assert torch.equal(t.result()[0], ref)


def test_gray_u8_inputs_through_the_fused_preprocessing(model_p):
    """raw 200x200 grayscale u8 images -> preprocess_gray kernel -> the same tokens as feeding the float tensor
    the oracle's restatement of the reference transform produces (inference_p.py:148-158 semantics)."""
    model_p.set_precision("bf16")
    tok = M.Tokenizer()
    u8 = O.synth_gray_u8(16, seed=77)
    x = O.preprocess_gray(u8)
    xg = M.preprocess_gray(u8.to(DEV))
    assert xg.shape == (16, 3, 224, 224) and torch.equal(xg.cpu(), x)
    want, _ = M.generate(model_p, xg, tok, max_len=30)
    got = list(M.generate_stream(model_p, [u8.pin_memory(), u8.to(DEV)], tok, max_len=30))
    assert torch.equal(got[0][0], want) and torch.equal(got[1][0], want)
