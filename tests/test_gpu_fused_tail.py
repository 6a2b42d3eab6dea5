"""GPU parity tests of the fused to_logits + sampling tail (mmg_logits_fused: no [rows, V] logits buffer) against the materialised path
(mmg_linear + mmg_logits_sample, itself pinned to the oracle in test_gpu_kernels.py): same token ids, same scores up to the summation order
of the softmax denominator, for libmmg's Philox keying, injected noise and the ATen stream; rows whose sampled threshold cannot work
(constant logits) take the in-call fallback; too many of them raise the overflow word and generate() repeats on the materialised path."""
import math
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
bf = torch.bfloat16


def ops():
    from muse_maskgit_pytorch_b200 import ops as _ops
    return _ops


def _case(R, V, K, seed, b=None):
    g = torch.Generator(device="cuda").manual_seed(seed)
    e = torch.randn((R, K), generator=g, device="cuda").to(bf)
    w = (torch.randn((V, K), generator=g, device="cuda") * K ** -0.5).to(bf)
    b = b or 1
    assert R % b == 0
    nm = R // b
    n = nm + 3
    mp = torch.stack([torch.sort(torch.randperm(n, generator=g, device="cuda")[:nm]).values for _ in range(b)]).int().contiguous()
    return e, w, b, n, nm, mp


def _run_both(e, w, b, n, nm, mp, k, temp, **kw):
    o = ops()
    R, V = e.shape[0], w.shape[0]
    ids_a = torch.full((b, n), V, dtype=torch.long, device="cuda"); sc_a = torch.full((b, n), -1e5, device="cuda")
    ids_b, sc_b = ids_a.clone(), sc_a.clone()
    lg = torch.empty((R, V), device="cuda")
    o.linear(e, w, lg)
    o.logits_sample(lg, mp, ids_a, sc_a, nm, k, temp, **kw)
    nbytes = o.logits_fused_workspace_bytes(R, V, e.shape[1], k)
    assert nbytes > 0
    ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
    status = torch.zeros((2,), dtype=torch.int32, device="cuda")
    o.logits_fused(e, w, mp, ids_b, sc_b, nm, k, temp, ws, status, rows_capacity=R, **kw)
    torch.cuda.synchronize()
    return ids_a, sc_a, ids_b, sc_b, status


@pytest.mark.parametrize("R,V,b", [(300, 65536, 1), (64, 65536, 2), (2048, 65536, 8), (128 * 67, 65536, 67), (700, 8192, 7), (96, 1024, 3), (260, 4096, 2)])
@pytest.mark.parametrize("temp", [1.0, 0.0])
def test_fused_equals_materialised_philox(R, V, b, temp):
    """libmmg Philox keying: identical sampled tokens; scores equal to 2e-6 (the fused path sums the softmax denominator per list segment)."""
    e, w, b, n, nm, mp = _case(R, V, 512 if V == 65536 else 128, seed=R + V, b=b)
    k = math.ceil(0.1 * V)
    ids_a, sc_a, ids_b, sc_b, status = _run_both(e, w, b, n, nm, mp, k, temp, seed=1234, step=3, row_offset=5 * n)
    assert torch.equal(ids_a, ids_b), int((ids_a != ids_b).sum())
    assert float((sc_a - sc_b).abs().max()) < 2e-6
    assert int(status[1]) == 0
    print(f"R={R} V={V}: fallback rows {int(status[0])}")


def test_fused_equals_materialised_injected_noise_and_aten():
    e, w, b, n, nm, mp = _case(240, 8192, 128, seed=7, b=4)
    k = math.ceil(0.1 * 8192)
    u = torch.rand((b, n, 8192), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    ids_a, sc_a, ids_b, sc_b, status = _run_both(e, w, b, n, nm, mp, k, 0.7, u=u)
    assert torch.equal(ids_a, ids_b) and float((sc_a - sc_b).abs().max()) < 2e-6
    off = torch.tensor([8], dtype=torch.int64, device="cuda"); seed_dev = torch.tensor([99], dtype=torch.int64, device="cuda")
    ids_a, sc_a, ids_b, sc_b, status = _run_both(e, w, b, n, nm, mp, k, 0.7, seed=0, seed_dev=seed_dev, aten=(16, off, 256 * 40))
    assert torch.equal(ids_a, ids_b) and float((sc_a - sc_b).abs().max()) < 2e-6


def test_fused_other_k_and_unsupported_shapes():
    """top-k thresholds other than 0.9; shapes the fused path declines (the caller then uses the materialised path)."""
    e, w, b, n, nm, mp = _case(256, 65536, 512, seed=11, b=2)
    o = ops()
    V = 65536
    for k in (math.ceil(0.05 * V), math.ceil(0.08 * V)):
        assert o.logits_fused_workspace_bytes(256, V, 512, k) > 0
        ids_a, sc_a, ids_b, sc_b, status = _run_both(e, w, b, n, nm, mp, k, 0.9, seed=5, step=1)
        assert torch.equal(ids_a, ids_b) and float((sc_a - sc_b).abs().max()) < 2e-6 and int(status[1]) == 0
    assert o.logits_fused_workspace_bytes(256, V, 512, math.ceil(0.2 * V)) == 0          # candidate lists would not fit: materialised path only
    assert o.logits_fused_workspace_bytes(256, 1000, 512, 100) == 0 and o.logits_fused_workspace_bytes(256, V, 100, 6554) == 0


def test_fused_fallback_rows_and_overflow_word():
    """Rows with constant logits (zero embedding): every logit ties, the candidate lists overflow, the row is redone from materialised logits
    inside the call (top-k keeps the lowest indices among equals -> same token as the materialised kernel).  More than 128 such rows in one
    call raise status[1]."""
    e, w, b, n, nm, mp = _case(384, 65536, 512, seed=13, b=3)
    k = math.ceil(0.1 * 65536)
    e[5] = 0; e[200] = 0; e[383] = 0
    ids_a, sc_a, ids_b, sc_b, status = _run_both(e, w, b, n, nm, mp, k, 1.0, seed=77, step=2)
    assert torch.equal(ids_a, ids_b) and float((sc_a - sc_b).abs().max()) < 2e-6
    assert int(status[0]) >= 3 and int(status[1]) == 0
    e[:200] = 0
    *_, status = _run_both(e, w, b, n, nm, mp, k, 1.0, seed=77, step=2)
    assert int(status[1]) == 1 and int(status[0]) >= 200


def test_generate_fused_tail_equals_materialised_and_overflow_rerun():
    """MaskGit.generate(): the fused tail (default) and the materialised tail give the same tokens and pixels, eager and under the CUDA graph;
    a model whose logits are constant (zero to_logits) overflows the fallback and is transparently re-run on the materialised path."""
    import muse_maskgit_pytorch_b200 as M
    from muse_maskgit_pytorch_b200 import t5
    t5.T5_CONFIGS["synth-128"] = {"d_model": 128}

    def build():
        tr = M.MaskGitTransformer(num_tokens=1024, dim=128, seq_len=16, depth=2, heads=2, t5_name="synth-128", precision="bf16")
        sd = util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128)
        full = tr.state_dict(); full.update(sd); tr.load_state_dict(full)
        vae = M.VQGanVAE(dim=16, layers=2, codebook_size=1024, precision="bf16")
        full = vae.state_dict(); full.update(util.vae_sd(16, 2, 1024, seed=12)); vae.load_state_dict(full)
        return M.MaskGit(image_size=16, transformer=tr.cuda(), vae=vae.cuda()).cuda()
    mg = build()
    te = util.text_embeds("g4.te", 3, 8, 128, 14).cuda()
    mg.transformer.encode_text = lambda texts: te[:len(texts)]
    mg.sampler_seed = 42
    outs = {}
    for fused in (True, False):
        for graph in (False, True):
            mg.use_fused_tail, mg.use_cuda_graph = fused, graph
            outs[(fused, graph)] = mg.generate(["a"] * 3, timesteps=8, return_ids=True)
    ref_img, ref_ids = outs[(False, False)]
    for key, (img, ids) in outs.items():
        assert torch.equal(ids, ref_ids), key
        assert float((img - ref_img).abs().max()) < 1e-6, key
    assert mg.last_fused_fallback_rows >= 0
    # degenerate model at V = 65536: all logits equal -> every row's lists overflow -> more fallback rows than the call holds -> overflow word
    # -> generate() repeats the call on the materialised path (same tokens as asking for that path directly)
    torch.manual_seed(5)
    tr = M.MaskGitTransformer(num_tokens=65536, dim=128, seq_len=16, depth=1, heads=2, t5_name="synth-128", precision="bf16")
    with torch.no_grad():
        tr.to_logits.weight.zero_()
    vae = M.VQGanVAE(dim=16, layers=2, codebook_size=65536, precision="bf16")
    mg2 = M.MaskGit(image_size=16, transformer=tr.cuda(), vae=vae.cuda()).cuda()
    b_big = 16                                            # 16 x 16 = 256 rows > 128 fallback rows on the first step
    te_big = te.repeat(6, 1, 1)
    mg2.transformer.encode_text = lambda texts: te_big[:len(texts)]
    mg2.sampler_seed = 42
    mg2.use_fused_tail = True
    img_f, ids_f = mg2.generate(["a"] * b_big, timesteps=4, return_ids=True)
    mg2.use_fused_tail = False
    img_m, ids_m = mg2.generate(["a"] * b_big, timesteps=4, return_ids=True)
    assert torch.equal(ids_f, ids_m) and torch.equal(img_f, img_m)
