"""GPU: a model on cuda:1 while the current device is cuda:0 (per-device launch state, device guards).  Needs two GPUs
(`gpurun --gpus 2`); skipped on a single-GPU box."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import mdcnet_b200 as M  # noqa: E402
from oracle import cases  # noqa: E402


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_model_on_second_device_with_first_device_current():
    torch.cuda.set_device(0)
    m0 = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda:0").set_precision("bf16")
    m1 = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda:1").set_precision("bf16")
    x = cases.images(5, seed=8)
    t0, c0 = m0.generate_tokens(x.to("cuda:0"), 12)
    t1, c1 = m1.generate_tokens(x.to("cuda:1"), 12)          # current device is still cuda:0: the engine guards the calls itself
    assert t1.device.index == 1 and torch.equal(t0.cpu(), t1.cpu()) and torch.equal(c0.cpu(), c1.cpu())
    p0 = m0.predict(x.to("cuda:0"), t0[:, :8].long()); p1 = m1.predict(x.to("cuda:1"), t1[:, :8].long())
    assert torch.equal(p0.cpu(), p1.cpu())
    # the batch pipeline on the second device (plans, graphs and streams of cuda:1 while cuda:0 is current)
    tok = M.Tokenizer()
    xs = [cases.images(5, seed=40 + i) for i in range(3)]
    got1 = list(M.generate_stream(m1, (v.to("cuda:1") for v in xs), tok, max_len=12))
    got0 = list(M.generate_stream(m0, (v.to("cuda:0") for v in xs), tok, max_len=12))
    assert all(torch.equal(a[0], b[0]) for a, b in zip(got0, got1))
    assert torch.equal(got1[0][0], M.generate(m1, xs[0].to("cuda:1"), tok, max_len=12)[0])
    assert torch.cuda.current_device() == 0
    boxes = torch.rand(3, 4, 4); boxes[..., 2:] += boxes[..., :2]
    assert torch.equal(torch.stack(M.calculate_batch_iou(boxes.to("cuda:1"), boxes.to("cuda:1"))).cpu(),
                       torch.stack(M.calculate_batch_iou(boxes.to("cuda:0"), boxes.to("cuda:0"))).cpu())
    # the C side refuses a context / device mismatch instead of launching on the wrong device
    L = M._lib
    a = torch.zeros(8, 8, device="cuda:1")
    with pytest.raises(L.MdcError):
        L.check(L.lib().mdc_layernorm(L.ctx(torch.device("cuda:1")), L.ptr(a), 8, L.ptr(a), L.ptr(a), 1e-5, L.ptr(a), 8, L.MDC_F32, 8, 8, L.stream_ptr()))
