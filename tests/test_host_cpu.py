"""CPU: host-side logic, the C-ABI library's export list, and the loud failure without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_the_header_declares():
    import mdcnet_b200 as M
    hdr = open(os.path.join(ROOT, "include", "mdc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mdc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(M._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mdc_b200.h but not exported"
    assert declared == set(M._lib.SIGNATURES), declared ^ set(M._lib.SIGNATURES)
    assert lib.mdc_abi_version() == int(re.search(r"#define MDC_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "mdc_b200.h")).read()).group(1))


def test_no_cpu_fallback():
    import mdcnet_b200 as M
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(M._lib.MdcError):
        M._lib.ctx()
    from oracle import cases
    model = cases.build_product_model("S", seed=0)
    with pytest.raises(M._lib.MdcError):
        model.predict(torch.zeros(1, 3, 224, 224), torch.tensor([[300]]))
    with pytest.raises(M._lib.MdcError):
        M.bbox_iou(torch.zeros(1, 4), torch.zeros(1, 4))
    # the C entry point itself refuses too
    lib = M._lib.lib()
    h = ctypes.c_void_p()
    assert lib.mdc_ctx_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.mdc_last_error()


def test_product_never_imports_the_oracle():
    pk = os.path.join(ROOT, "mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("mdc_oracle", "oracle") or "import" not in "".join(
                    l for l in src.splitlines() if "oracle" in l), f


def test_state_dict_keys_follow_the_reference_layout():
    from oracle import cases
    m = cases.build_product_model("S", seed=0)
    keys = set(m.state_dict())
    for k in ["encoder.model.cls_token", "encoder.model.pos_embed", "encoder.model.patch_embed.proj.weight",
              "encoder.model.blocks.11.attn.qkv.bias", "encoder.model.blocks.0.ls1.gamma", "encoder.model.blocks.3.mlp.fc2.weight",
              "encoder.model.norm.bias", "decoder.decoder_pos_embed", "decoder.encoder_pos_embed", "decoder.embedding.weight",
              "decoder.output.bias", "decoder.decoder.layers.1.self_attn.in_proj_weight",
              "decoder.decoder.layers.0.multihead_attn.out_proj.bias", "decoder.decoder.layers.1.linear2.weight",
              "decoder.decoder.layers.0.norm3.weight"]:
        assert k in keys, k
    ax = cases.build_product_model("S", seed=0, axial=True)
    assert {"decoder.axial_attention.to_qkv.weight", "decoder.axial_attention.to_out.weight",
            "decoder.axial_attention.to_out.bias"} <= set(ax.state_dict())
    assert "decoder.axial_attention.to_qkv.bias" not in ax.state_dict()


def test_page_allocator_and_table():
    from mdcnet_b200.kvcache import PageAllocator, build_page_table, pages_for
    assert pages_for(99, 16) == 7 and pages_for(16, 16) == 1 and pages_for(17, 16) == 2
    a = PageAllocator(4 * 7)
    t = build_page_table(a, 4, 99, 16, interleave=True)
    flat = [p for row in t for p in row]
    assert sorted(flat) == list(range(28)) and a.n_free == 0
    assert t[0][1] - t[0][0] == 4           # interleaved: a sequence's pages are not contiguous
    with pytest.raises(MemoryError):
        a.alloc(9, 1)
    assert a.release(2) == 7 and a.n_free == 7
    assert len(a.alloc(9, 7)) == 7


def test_shard_bounds_and_packing():
    from mdcnet_b200 import parallel as P
    for B, W in [(512, 8), (512, 4), (64, 1), (10, 4), (3, 8)]:
        spans = [P.shard_bounds(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        sizes = [e - s for s, e in spans]
        assert max(sizes) - min(sizes) <= 1
    tok = torch.randint(0, 305, (5, 100), dtype=torch.int32)
    conf = torch.rand(5, 25)
    boxes = torch.rand(5, 3, 4); mi = torch.rand(5, 3)
    t2, c2, b2, m2 = P.unpack_results(P.pack_results(tok, conf, boxes, mi), 100, 25, 3)
    assert torch.equal(t2, tok) and torch.equal(c2, conf) and torch.equal(b2, boxes) and torch.equal(m2, mi)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MDC_ROOT"])
from mdcnet_b200 import parallel as P
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MDC_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world, B = dist.get_rank(), dist.get_world_size(), 7
s, e = P.shard_bounds(B, rank, world)
g = torch.Generator().manual_seed(0)
tok = torch.randint(0, 305, (B, 11), generator=g, dtype=torch.int32); conf = torch.rand(B, 3, generator=g)
full = P.all_gather_results(P.pack_results(tok[s:e], conf[s:e]), B)
t2, c2, _, _ = P.unpack_results(full, 11, 3)
assert torch.equal(t2, tok) and torch.equal(c2, conf), rank
dist.destroy_process_group()
print("ok", rank)
"""


def test_all_gather_results_world2_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MDC_PORT=str(port), MDC_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, "-c", _GLOO_WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()


def test_tokenizer_decode_side_has_no_cpu_path():
    """Token -> box decoding is a GPU kernel (csrc/tokens.cu); on a host without CUDA it must fail loudly, not fall back."""
    import mdcnet_b200 as M
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the -m gpu tests")
    tk = M.Tokenizer(num_bins=224, width=224, height=224)
    with pytest.raises(Exception):
        tk.decode_bboxes(torch.tensor([[300, 303, 270, 304, 259, 10, 20, 110, 220, 301]]))


def test_generate_rejects_non_b200_model():
    import mdcnet_b200 as M
    with pytest.raises(TypeError):
        M.generate(torch.nn.Linear(2, 2), torch.zeros(1, 3, 224, 224), M.Tokenizer(), max_len=3)
