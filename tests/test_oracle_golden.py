"""CPU: the oracle (oracle/muse_oracle.py) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle (task section 3)."""
import json
import os
import numpy as np
import pytest
import torch

from oracle import muse_oracle as O
from oracle import synth
from tests import util

torch.set_grad_enabled(False)
CFG = dict(heads=2, depth=2)


def test_known_answers():
    ka = json.load(open(os.path.join(util.GOLDEN, "known_answers.json")))
    assert O.mask_schedule(256, 18) == ka["sched_256_18"] == [256, 254, 251, 246, 238, 229, 217, 204, 189, 172, 154, 134, 114, 92, 70, 47, 23, 1]
    assert O.mask_schedule(1024, 18) == ka["sched_1024_18"]
    assert O.mask_schedule(16, 6) == ka["sched_16_6"] and O.mask_schedule(64, 8) == ka["sched_64_8"]
    for v, k in ka["topk"].items():
        assert O.top_k_count(int(v), 0.9) == k
    assert O.top_k_count(65536, 0.9) == 6554


def test_vae_c1_config():
    g = util.golden("vae_c1")
    sd = util.vae_sd(64, 4, 512, seed=11)
    img = torch.from_numpy(synth.uniform("c1.img", (4, 3, 32, 32), 11))
    fmap = O.vae_encode_fmap(sd, img)
    assert torch.allclose(fmap, g["fmap"], atol=1e-5, rtol=1e-5)
    q, ids = O.vae_encode(sd, img)
    assert torch.equal(ids, g["ids"]) and ids.min() >= 0 and ids.max() < 512
    assert torch.allclose(q, g["fmap_q"], atol=1e-5)
    rec = O.vae_decode_from_ids(sd, ids, 9)
    assert torch.allclose(rec, g["recon"], atol=1e-5, rtol=1e-5)


def test_vae_small():
    g = util.golden("vae_small")
    sd = util.vae_sd(16, 2, 1024, seed=12)
    img = torch.from_numpy(synth.uniform("g2.img", (2, 3, 16, 16), 12))
    q, ids = O.vae_encode(sd, img)
    assert torch.equal(ids, g["ids"])
    assert torch.allclose(O.vae_decode_from_ids(sd, ids, 10), g["recon"], atol=1e-5)


def test_lfq_is_l2_argmin_over_sign_codebook():
    """Pin for the un-vendored LFQ: ids == argmin_k ||x - c_k||^2 over the explicit {+-1}^d codebook
    (MSB-first bit order), including exact-zero coordinates (-> bit 0 -> lower index)."""
    d = 6
    x = torch.from_numpy(synth.dyadic("lfq.x", (200, d), bits=3, span=2.0))
    x[::7, 2] = 0.
    codes = torch.tensor([[1. if (k >> (d - 1 - i)) & 1 else -1. for i in range(d)] for k in range(2 ** d)])
    _, ids = O.lfq_quantize({}, x.t().reshape(1, d, 200, 1).permute(0, 1, 2, 3))
    brute = O.vq_l2_argmin(x, codes)
    assert torch.equal(ids.reshape(-1), brute)


@pytest.mark.parametrize("tag,d_text", [("tr_small", 128), ("tr_small_proj", 96), ("tr_small_flash", 128)])
def test_transformer_forward(tag, d_text):
    g = util.golden(tag)
    sd = util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=d_text)
    te = util.text_embeds("g3.te", 3, 8, d_text, 13)
    ids = g["ids"]
    cfg_logits, embed = O.forward_with_cond_scale(sd, CFG, ids, te, cond_scale=3.)
    tol = 2e-5 if tag != "tr_small_flash" else 1e-4     # flash stand-in: tiled online softmax, same math
    assert torch.allclose(embed, g["embed"], atol=tol, rtol=1e-4)
    assert torch.allclose(cfg_logits, g["logits_cfg"], atol=5 * tol, rtol=1e-4)
    if "logits_null" in g:
        assert torch.allclose(O.transformer_forward(sd, CFG, ids, te, drop_text=True), g["logits_null"], atol=tol, rtol=1e-4)


@pytest.mark.parametrize("T", [6, 18])
def test_generate_small(T):
    g = util.golden(f"gen_small_T{T}")
    torch.manual_seed(777)
    if not torch.equal(torch.zeros(7).uniform_(0, 1), g["u_probe"]):
        pytest.skip("torch CPU generator differs from the one that made the fixture")
    sd = util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128)
    vsd = util.vae_sd(16, 2, 1024, seed=12)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    images, ids = O.generate(sd, CFG, vsd, 10, te, 4, util.torch_noise_fn(777), timesteps=T)
    assert torch.equal(ids.view(3, 4, 4), g["ids"])
    assert torch.allclose(images, g["images"], atol=1e-5)


def test_generate_superres_small():
    g = util.golden("gen_superres_small")
    sd = util.transformer_sd(1024, 128, 64, 2, 2, seed=15, text_dim=128)
    vsd = util.vae_sd(16, 2, 1024, seed=12)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)[:2]
    cond = torch.from_numpy(synth.uniform("g5.cond", (2, 3, 16, 16), 15))
    _, cond_ids = O.vae_encode(vsd, cond)
    assert torch.equal(cond_ids, g["cond_ids"])
    images, ids = O.generate(sd, CFG, vsd, 10, te, 8, util.torch_noise_fn(778), cond_images=cond, timesteps=8)
    assert torch.equal(ids.view(2, 8, 8), g["ids"])
    assert torch.allclose(images, g["images"], atol=1e-5)


def _branch_case(name):
    """(golden, oracle kwargs, torch seed) of the optional-branch fixtures G6-G8 (make_golden.py)."""
    cfg_sc = dict(CFG, self_cond=True)
    if name == "gen_selfcond_small":
        return util.transformer_sd(1024, 128, 16, 2, 2, seed=21, text_dim=128), cfg_sc, {}, 779
    if name == "gen_critic_small":
        crit = dict(kind="token", sd=util.critic_sd(1024, 128, 16, 1, 2, seed=22, text_dim=128), cfg=dict(heads=2, depth=1))
        return util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128), CFG, dict(critic=crit, critic_noise_scale=0.7), 780
    if name == "gen_critic_forced_off_small":
        return util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128), CFG, {}, 780
    if name == "gen_remask_prev_small":
        return util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128), CFG, dict(can_remask_prev_masked=True), 782
    sd = util.transformer_sd(1024, 128, 16, 2, 2, seed=21, text_dim=128)
    w, b = util.self_critic_head()
    return sd, cfg_sc, dict(critic=dict(kind="self", sd=sd, cfg=cfg_sc, w_pred=w, b_pred=b)), 781


@pytest.mark.parametrize("name", ["gen_selfcond_small", "gen_critic_small", "gen_critic_forced_off_small", "gen_selfcritic_small",
                                  "gen_remask_prev_small"])
def test_generate_optional_branches(name):
    """self-conditioning feedback, TokenCritic and SelfCritic scoring against the unmodified reference."""
    g = util.golden(name)
    sd, cfg, kw, seed = _branch_case(name)
    vsd = util.vae_sd(16, 2, 1024, seed=12)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    images, ids = O.generate(sd, cfg, vsd, 10, te, 4, util.torch_noise_fn(seed), timesteps=8, **kw)
    assert torch.equal(ids.view(3, 4, 4), g["ids"])
    assert torch.allclose(images, g["images"], atol=1e-5)
    if kw or cfg.get("self_cond"):       # the fixture is discriminative: dropping the branch changes the tokens
        plain, _ = O.generate_ids(sd, dict(cfg, self_cond=False), te, 16, 1024, util.torch_noise_fn(seed), timesteps=8), None
        assert not torch.equal(plain.view(3, 4, 4), g["ids"])


def test_philox_known_answer():
    """Random123 known-answer vector for Philox4x32-10 (zero counter, zero key -> 6627e8d5 ...): pins oracle/philox.py,
    against which the in-kernel generator is tested on the GPU."""
    from oracle import philox
    u = philox.uniform(0, 0, 0, 1)
    assert int(round(float(u[0]) * (1 << 24))) == 0x6627e8d5 >> 8
    a, b = philox.uniform(5, 1, 2, 64), philox.uniform(5, 1, 3, 64)
    assert not np.array_equal(a, b) and a.min() >= 0 and a.max() < 1


def test_aten_uniform_model_known_answers():
    """oracle/philox.aten_uniform restates ATen's CUDA uniform_ stream (pinned against torch.cuda itself on the GPU box,
    tests/test_gpu_aten_rng.py).  Offline regression: the first values of `torch.manual_seed(0); torch.rand(3, device="cuda")`
    and the launch geometry / generator-offset bookkeeping for the C3 noise tensor on a 148-SM device."""
    from oracle import philox
    u = philox.aten_uniform(0, 0, 3, philox.aten_stride(3, 148, 2048))
    assert [f"{x:.4f}" for x in u] == ["0.3990", "0.5167", "0.0249"]
    numel = 64 * 256 * 65536
    stride = philox.aten_stride(numel, 148, 2048)
    assert stride == 148 * 8 * 256 and philox.aten_offset_increment(numel, stride) == ((numel - 1) // (stride * 4) + 1) * 4 == 3544
    # element li = t + stride * (4 j + ii) is word ii of Philox(counter = offset / 4 + j, subsequence = t): spot-check the indexing
    li = np.array([5, stride + 5, 4 * stride + 5], dtype=np.uint64)
    a = philox.aten_uniform(7, 8, numel, stride, index=li)
    w = philox.philox4(np.array([2, 2, 3], dtype=np.uint64), 0, np.array([5, 5, 5], dtype=np.uint64), 0, 7, 0)
    want = [float(np.float32(w[0][0]) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)),
            float(np.float32(w[1][1]) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)),
            float(np.float32(w[0][2]) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33))]
    assert [float(x) for x in a] == want


def _noise_stream(seed):
    """One torch CPU stream across BOTH generate() calls of the cascade (the reference seeds once, muse_maskgit_pytorch.py:758-791)."""
    torch.manual_seed(seed)
    return lambda step, shape: torch.zeros(shape).uniform_(0, 1)


def test_muse_cascade_small():
    """Muse.forward = base generate -> super-res generate conditioned on the low-res images (golden G10, unmodified reference)."""
    g = util.golden("muse_small")
    vsd = util.vae_sd(16, 2, 1024, seed=12)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)[:2]
    noise = _noise_stream(783)
    low, base_ids = O.generate(util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128), CFG, vsd, 10, te, 4, noise, timesteps=6)
    assert torch.equal(base_ids.view(2, 4, 4), g["base_ids"])
    assert torch.allclose(low, g["lowres"], atol=1e-5)
    sup, sr_ids = O.generate(util.transformer_sd(1024, 128, 64, 2, 2, seed=15, text_dim=128), CFG, vsd, 10, te, 8, noise, cond_images=low, timesteps=8)
    assert torch.equal(sr_ids.view(2, 8, 8), g["superres_ids"])
    assert torch.allclose(sup, g["superres"], atol=1e-5)


def test_reference_written_checkpoints():
    """Checkpoints saved by the reference's own save() (G11): the oracle on those weights reproduces the reference's cascade, and the
    files carry the training-only tensors (discr.*) the drop-in has to skip."""
    import os
    g = util.golden("ckpt_muse")
    load = lambda n: torch.load(os.path.join(util.GOLDEN, n), map_location="cpu")
    base, sr, vae = load("ckpt_base.pt"), load("ckpt_superres.pt"), load("ckpt_vae.pt")
    assert any(k.startswith("discr.") for k in vae) and any(k.startswith("cond_vae.discr.") for k in sr)
    sub = lambda sd, pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    cfg = dict(heads=1, depth=1)
    te = util.text_embeds("g11.te", 2, 8, 64, 31)
    img = torch.from_numpy(synth.uniform("g11.img", (2, 3, 16, 16), 31))
    _, vids = O.vae_encode(vae, img)
    assert torch.equal(vids, g["vae_ids"])
    assert torch.allclose(O.vae_decode_from_ids(vae, vids, 8), g["vae_recon"], atol=1e-5)
    noise = _noise_stream(901)
    low, base_ids = O.generate(sub(base, "transformer."), cfg, sub(base, "vae."), 8, te, 4, noise, timesteps=4)
    assert torch.equal(base_ids.view(2, 4, 4), g["base_ids"]) and torch.allclose(low, g["lowres"], atol=1e-5)
    sup, sr_ids = O.generate(sub(sr, "transformer."), cfg, sub(sr, "vae."), 8, te, 8, noise, cond_images=low, sd_cond_vae=sub(sr, "cond_vae."), timesteps=4)
    assert torch.equal(sr_ids.view(2, 8, 8), g["superres_ids"]) and torch.allclose(sup, g["superres"], atol=1e-5)


def test_split3_six_terms_carry_the_fp32_product():
    """oracle/split3.py (the restatement of mmg_split3 the GPU tests check the kernel against): hi + mid + lo reconstructs fp32 to 2^-23 relative,
    every term is a bf16 value, and the six cross terms of the 6K-wide product reproduce the fp64 product of the fp32 operands to ~2^-23 of
    sum |a||w| — two orders below one bf16-operand product."""
    from oracle import split3 as S
    rng = np.random.default_rng(7)
    a = (rng.standard_normal((96, 192)) * np.exp(rng.standard_normal((96, 192)) * 3)).astype(np.float32)
    w = (rng.standard_normal((80, 192)) * 192 ** -0.5).astype(np.float32)
    t = S.terms(a)
    for v in t.values():
        assert np.array_equal(S.bf16_round(v), v)                                  # each term is exactly representable in bf16
    assert np.all(np.abs(t["hi"].astype(np.float64) + t["mid"] + t["lo"] - a) <= np.abs(a) * 2.0 ** -23)
    assert S.split3(a, 0).shape == (96, 6 * 192) and S.split3(w, 1).shape == (80, 6 * 192)
    exact = a.astype(np.float64) @ w.astype(np.float64).T
    scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T
    err6 = np.abs(S.product(a, w) - exact) / scale
    err1 = np.abs(S.bf16_round(a).astype(np.float64) @ S.bf16_round(w).astype(np.float64).T - exact) / scale
    assert err6.max() < 2.0 ** -21 and err1.max() > 50 * err6.max()


def test_gelu_polynomial_of_the_tensor_core_epilogue_is_exact_erf_gelu_to_1e6():
    """The GEGLU epilogue's GELU (csrc/mmg_common.cuh: gelu_fast = max(x, 0) - 0.5 |x| 2^q(min(|x|, 6)), q a degree-6 polynomial for
    log2 erfc(a / sqrt 2)) against the reference's exact-erf F.gelu (muse_maskgit_pytorch.py:76-77), evaluated here in fp32 with the
    coefficients READ FROM THE KERNEL SOURCE: max abs error < 1e-6 over [-12, 12], relative error < 2e-4 wherever |gelu| > 1e-3."""
    import re
    from scipy.special import erf
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "muse_maskgit_pytorch_b200", "csrc", "mmg_common.cuh")).read()
    body = src[src.index("float gelu_fast(float x)"):]
    body = body[:body.index("return fmaf(-0.5f * a, e, fmaxf(x, 0.0f));")]
    c = [np.float32(v) for v in re.findall(r"(-?\d+\.\d+(?:e-?\d+)?)f", body)]
    assert len(c) == 7 and c[0] == np.float32(6.0), c                                # clamp, then c6, c5, c4, c3, c2, c1 in Horner order
    x = np.linspace(-12, 12, 480001).astype(np.float32)
    a = np.abs(x); aq = np.minimum(a, c[0])
    q = (c[1] * aq + c[2]).astype(np.float32)
    for ci in c[3:]:
        q = (q * aq + ci).astype(np.float32)
    e = np.exp2((q * aq).astype(np.float64)).astype(np.float32)
    got = (np.maximum(x, np.float32(0)) - np.float32(0.5) * a * e).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1.0 + erf(x.astype(np.float64) / np.sqrt(2.0)))
    err = np.abs(got - ref)
    assert err.max() < 1e-6, err.max()
    big = np.abs(ref) > 1e-3
    assert (err[big] / np.abs(ref[big])).max() < 2e-4


def test_fused_tail_selection_equals_topk_gumbel_argmax():
    """oracle/fused_tail.py: the candidate-list procedure of the fused sampling tail (threshold superset, perturbed argmax, exact-rank test,
    exclusion + repeat) picks the reference's token on rows with ties, flat rows (many repeats) and peaked rows, for thresholds anywhere between
    'barely k candidates' and 'the whole row'."""
    from oracle import fused_tail as FT
    rng = np.random.default_rng(11)
    V, k = 2048, 205
    repeats = 0
    for case in range(60):
        kind = case % 4
        if kind == 0:
            logits = rng.standard_normal(V).astype(np.float32)
        elif kind == 1:
            logits = np.round(rng.standard_normal(V) * 2).astype(np.float32) / 2          # heavy ties, also across the k-th value
        elif kind == 2:
            logits = (rng.standard_normal(V) * 0.01).astype(np.float32)                     # flat: the winner often falls outside the top-k
        else:
            logits = (rng.standard_normal(V) * 6).astype(np.float32)                        # peaked
        u = rng.random(V).astype(np.float32)
        g = (-np.log(np.maximum(-np.log(np.maximum(u, 1e-20)), 1e-20))).astype(np.float32)
        T = [1.0, 0.3, 2.5][case % 3]
        want = FT.reference_choice(logits, g, k, T)
        srt = np.sort(logits)[::-1]
        for n_cand in (k, k + 1, k + 40, 2 * k, V):
            thr = srt[n_cand - 1]
            got, passes = FT.fused_choice(logits, g, k, T, thr, max_excl=V)
            assert got == want, (case, n_cand, got, want)
            repeats += passes - 1
        assert FT.fused_choice(logits, g, k, T, np.float32(np.inf))[0] is None              # fewer than k candidates: materialised path
    assert repeats > 0                                                                       # the exclusion path was exercised
