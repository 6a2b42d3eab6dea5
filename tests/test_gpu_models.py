"""GPU parity tests at the class boundary: the drop-in VQGanVAE / MaskGitTransformer / MaskGit against the golden vectors
produced by the unmodified reference (tests/golden) and against the CPU oracle.

Stated tolerances (north_star: ids bit-identical, logits/pixels within a stated fp tolerance):
  precision="fp32": logits max-abs <= 2e-4, pixels max-abs <= 1e-4, VQ ids and generated token ids identical to the reference;
  precision="bf16": logits rel-L2 <= 2e-2 (the reference's own bf16-vs-fp32 figure is 1.3e-2), pixels max-abs <= 5e-2;
                    token ids are compared step by step (teacher forced) with a flip-rate bound, since no bf16 implementation
                    can be token-identical to an fp32 reference.
"""
import numpy as np
import pytest
import torch

from oracle import synth, muse_oracle as O
from tests import util

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
CFG = dict(heads=2, depth=2)


def M():
    import muse_maskgit_pytorch_b200 as m
    from muse_maskgit_pytorch_b200 import t5
    for d in (96, 128, 512):
        t5.T5_CONFIGS[f"synth-{d}"] = {"d_model": d}
    return m


def load(module, sd):
    full = module.state_dict()
    for k, v in sd.items():
        assert k in full, k
        full[k] = v
    module.load_state_dict(full)
    return module.cuda()


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def small_transformer(precision, d_text=128, seq_len=16, seed=13):
    tr = M().MaskGitTransformer(num_tokens=1024, dim=128, seq_len=seq_len, depth=2, heads=2, t5_name=f"synth-{d_text}", precision=precision)
    return load(tr, util.transformer_sd(1024, 128, seq_len, 2, 2, seed=seed, text_dim=d_text))


def small_vae(precision):
    return load(M().VQGanVAE(dim=16, layers=2, codebook_size=1024, precision=precision), util.vae_sd(16, 2, 1024, seed=12))


# ------------------------------------------------------------------------------------------------ transformer forward
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,d_text", [("tr_small", 128), ("tr_small_proj", 96)])
def test_transformer_forward_vs_reference_golden(precision, tag, d_text):
    g = util.golden(tag)
    tr = small_transformer(precision, d_text)
    te = util.text_embeds("g3.te", 3, 8, d_text, 13).cuda()
    ids = g["ids"].cuda()
    logits, embed = tr.forward_with_cond_scale(ids, text_embeds=te, cond_scale=3., return_embed=True)
    if precision == "fp32":
        assert (embed.cpu() - g["embed"]).abs().max() < 5e-5
        assert (logits.cpu() - g["logits_cfg"]).abs().max() < 2e-4
    else:
        assert rel_l2(embed, g["embed"]) < 1.5e-2, rel_l2(embed, g["embed"])
        assert rel_l2(logits, g["logits_cfg"]) < 2e-2, rel_l2(logits, g["logits_cfg"])
    if "logits_null" in g:
        null = tr(ids, text_embeds=te, cond_drop_prob=1.)
        cond = tr(ids, text_embeds=te)
        if precision == "fp32":
            assert (null.cpu() - g["logits_null"]).abs().max() < 2e-4 and (cond.cpu() - g["logits_cond"]).abs().max() < 2e-4
        else:
            assert rel_l2(null, g["logits_null"]) < 2e-2 and rel_l2(cond, g["logits_cond"]) < 2e-2


def test_transformer_c2_shape_bf16_vs_oracle():
    """config C2 geometry (dim 512, seq 256, 8 heads) at depth 2 / V=8192 so the CPU oracle stays fast."""
    m = M()
    tr = m.MaskGitTransformer(num_tokens=8192, dim=512, seq_len=256, depth=2, heads=8, t5_name="synth-512", precision="bf16")
    sd = util.transformer_sd(8192, 512, 256, 2, 8, seed=21, text_dim=512)
    load(tr, sd)
    te = util.text_embeds("c2.te", 2, 32, 512, 21)
    ids = torch.from_numpy((synth.uniform("c2.ids", (2, 256), 21) * 8193).astype(np.int64))
    ref, _ = O.forward_with_cond_scale(sd, dict(heads=8, depth=2), ids, te, cond_scale=3.)
    got = tr.forward_with_cond_scale(ids.cuda(), text_embeds=te.cuda(), cond_scale=3.)
    r = rel_l2(got, ref)
    agree = float((got.cpu().argmax(-1) == ref.argmax(-1)).float().mean())
    print(f"C2-shape bf16 logits rel-L2 {r:.3e}, argmax agreement {agree:.3f}")
    assert r < 2e-2 and agree > 0.9


# ------------------------------------------------------------------------------------------------ VAE
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vae_c1_config_vs_reference_golden(precision):
    """config C1: VQGanVAE dim=64 codebook=512 encode -> VQ -> decode on 4x3x32x32."""
    g = util.golden("vae_c1")
    vae = load(M().VQGanVAE(dim=64, codebook_size=512, precision=precision), util.vae_sd(64, 4, 512, seed=11))
    img = torch.from_numpy(synth.uniform("c1.img", (4, 3, 32, 32), 11)).cuda()
    fq, ids, aux = vae.encode(img)
    assert ids.shape == (4, 2, 2) and int(ids.min()) >= 0 and int(ids.max()) < 512 and float(aux) == 0.
    rec_from_ref_ids = vae.decode_from_ids(g["ids"].cuda())
    rec = vae(img)
    if precision == "fp32":
        assert torch.equal(ids.cpu(), g["ids"])
        assert (fq.cpu() - g["fmap_q"]).abs().max() < 1e-5
        assert (rec_from_ref_ids.cpu() - g["recon"]).abs().max() < 1e-4
        assert (rec.cpu() - g["recon"]).abs().max() < 1e-4
    else:
        proj = g["proj"].reshape(-1, 9)
        flips = (ids.cpu().reshape(-1, 1) >> torch.arange(8, -1, -1)) & 1 != (proj > 0).long()
        # a bit may only flip where the fp32 projection is within bf16 noise of zero
        assert (proj.abs()[flips] < 0.05).all(), proj.abs()[flips].max()
        assert (rec_from_ref_ids.cpu() - g["recon"]).abs().max() < 5e-2
    assert torch.equal(vae.decode_from_ids(ids).cpu(), rec.cpu())           # vae(x) == decode_from_ids(encode(x).ids)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vae_small_vs_reference_golden(precision):
    g = util.golden("vae_small")
    vae = small_vae(precision)
    img = torch.from_numpy(synth.uniform("g2.img", (2, 3, 16, 16), 12)).cuda()
    ids = vae.encode_ids(img)
    rec = vae.decode_from_ids(g["ids"].cuda())
    if precision == "fp32":
        assert torch.equal(ids.cpu(), g["ids"])
        assert (rec.cpu() - g["recon"]).abs().max() < 1e-4
    else:
        assert (rec.cpu() - g["recon"]).abs().max() < 5e-2


def test_vae_explicit_codebook_path():
    """lookup_free_quantization=False (the path the reference intends but crashes on, defect B1): ids == L2 argmin."""
    vae = M().VQGanVAE(dim=16, layers=2, codebook_size=256, lookup_free_quantization=False, precision="fp32")
    sd = {k: v for k, v in util.vae_sd(16, 2, 1024, seed=12).items() if not k.startswith("quantizer.")}
    load(vae, sd)
    img = torch.from_numpy(synth.uniform("g2.img", (2, 3, 16, 16), 12)).cuda()
    fq, ids, _ = vae.encode(img)
    fmap = O.vae_encode_fmap(sd, img.cpu()).permute(0, 2, 3, 1).reshape(-1, 32)
    ref = O.vq_l2_argmin(fmap, vae.quantizer.embed.detach().cpu())
    d = torch.cdist(fmap, vae.quantizer.embed.detach().cpu())
    top2 = d.topk(2, dim=-1, largest=False).values
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4
    assert torch.equal(ids.cpu().reshape(-1)[clear], ref[clear])
    assert vae.decode_from_ids(ids).shape == (2, 3, 16, 16)


def test_vae_explicit_codebook_bf16_argmin_sees_fp32_codebook():
    """bf16 precision, explicit codebook at a tensor-core shape (D = 64, 1024 codes): the nearest-code search runs against the fp32 codebook
    (three bf16 terms along K), so for the bf16 tokens the kernel is given the ids equal the fp32 L2 argmin — also where rounding the codebook to
    bf16 would have picked another code."""
    torch.manual_seed(5)
    vae = M().VQGanVAE(dim=64, layers=1, codebook_size=1024, lookup_free_quantization=False, precision="bf16").cuda()
    cb = vae.quantizer.embed.detach().float().cpu()
    assert cb.shape == (1024, 64)
    g = torch.Generator().manual_seed(9)
    x = (cb[torch.randint(0, 1024, (4096,), generator=g)] + 0.3 * torch.randn((4096, 64), generator=g)).to(torch.bfloat16)
    ids = vae._quantize_nhwc(x.cuda().contiguous()).cpu()
    ref = O.vq_l2_argmin(x.float(), cb)
    d = torch.cdist(x.float().double(), cb.double())
    top2 = d.topk(2, dim=-1, largest=False).values
    clear = (top2[:, 1] - top2[:, 0]) > 1e-5                      # fp32 summation noise of the reference argmin itself
    assert torch.equal(ids[clear], ref[clear])
    rounded = O.vq_l2_argmin(x.float(), cb.to(torch.bfloat16).float())
    print(f"explicit codebook, bf16 tokens: {int((rounded != ref).sum())} of 4096 tokens would change code with a bf16-rounded codebook; "
          f"{int((ids != ref).sum())} differ from the fp32 argmin ({int((~clear).sum())} near-ties)")


# ------------------------------------------------------------------------------------------------ generate
def make_maskgit(precision, superres=False):
    m = M()
    if not superres:
        return m.MaskGit(image_size=16, transformer=small_transformer(precision), vae=small_vae(precision)).cuda()
    return m.MaskGit(image_size=32, transformer=small_transformer(precision, seq_len=64, seed=15), vae=small_vae(precision), cond_image_size=16).cuda()


@pytest.mark.parametrize("T", [6, 18])
def test_generate_fp32_token_identical_to_reference(T):
    """End-to-end MaskGit.generate in parity precision with the reference's own noise stream injected:
    token ids bit-identical to the unmodified reference, pixels within 1e-4."""
    g = util.golden(f"gen_small_T{T}")
    torch.manual_seed(777)
    if not torch.equal(torch.zeros(7).uniform_(0, 1), g["u_probe"]):
        pytest.skip("torch CPU generator differs from the one that made the fixture")
    mg = make_maskgit("fp32")
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    mg.transformer.encode_text = lambda texts: te
    mg.sampler_noise_fn = util.torch_noise_fn(777)
    images, ids = mg.generate(texts=["a"] * 3, timesteps=T, return_ids=True)
    assert torch.equal(ids.cpu(), g["ids"]), (ids.cpu() != g["ids"]).sum()
    assert (images.cpu() - g["images"]).abs().max() < 1e-4


def test_generate_superres_fp32_token_identical_to_reference():
    g = util.golden("gen_superres_small")
    mg = make_maskgit("fp32", superres=True)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)[:2]
    mg.transformer.encode_text = lambda texts: te
    mg.sampler_noise_fn = util.torch_noise_fn(778)
    cond = torch.from_numpy(synth.uniform("g5.cond", (2, 3, 16, 16), 15)).cuda()
    images, ids = mg.generate(texts=["a"] * 2, cond_images=cond, timesteps=8, return_ids=True)
    assert torch.equal(ids.cpu(), g["ids"]), (ids.cpu() != g["ids"]).sum()
    assert (images.cpu() - g["images"]).abs().max() < 1e-4


def test_generate_bf16_teacher_forced_flip_rate():
    """bf16 fast path, step by step against the fp32 oracle trace: feed the oracle's ids/scores into each step and count
    sampled-token flips; bound = 8 % (the reference's own bf16-vs-fp32 argmax disagreement is 3-5 %)."""
    from muse_maskgit_pytorch_b200 import ops
    mg = make_maskgit("bf16")
    tr = mg.transformer
    sd = util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    trace = []
    O.generate_ids(sd, CFG, te, 16, 1024, util.torch_noise_fn(777), timesteps=18, trace=trace)
    ctx = tr._prepare_context(te.cuda(), None, [False, True])
    P = tr._packed()
    flips = total = 0
    for st in trace:
        ids_in = st["ids_in"].cuda()
        b, n = ids_in.shape
        nm = st["num_masked"]
        mp = torch.stack([torch.nonzero(ids_in[i] == 1024).flatten() for i in range(b)]).int().contiguous()
        x = tr._run_blocks(ids_in, ctx, 2)
        e = torch.empty((b * nm, 128), device="cuda", dtype=P["adt"])
        ops.final_embed(x[:b * n], x[b * n:], P["gf"], mp, e, b, n, nm, 3.0)
        lg = torch.empty((b * nm, 1024), device="cuda")
        ops.linear(e, P["wlog"], lg)
        ids = ids_in.clone(); sc = torch.full((b, n), -1e5, device="cuda")
        ops.logits_sample(lg, mp, ids, sc, nm, 103, float(st["temperature"]), u=st["u"].cuda().contiguous())
        is_mask = st["ids_in"] == 1024
        flips += int((ids.cpu()[is_mask] != st["ids_out"][is_mask]).sum()); total += int(is_mask.sum())
    print(f"bf16 teacher-forced flip rate {flips}/{total}")
    assert flips / total < 0.08


def test_generate_bf16_philox_shapes_and_determinism():
    mg = make_maskgit("bf16")
    te = util.text_embeds("g4.te", 3, 8, 128, 14).cuda()
    mg.transformer.encode_text = lambda texts: te[:len(texts)]
    mg.sampler_seed = 42
    a, ida = mg.generate(texts=["a"] * 3, timesteps=8, return_ids=True)
    b_, idb = mg.generate(texts=["a"] * 3, timesteps=8, return_ids=True)
    assert a.shape == (3, 3, 16, 16) and a.dtype == torch.float32 and torch.isfinite(a).all()
    assert torch.equal(ida, idb) and torch.equal(a, b_)
    assert int(ida.min()) >= 0 and int(ida.max()) < 1024
    # batch-shard invariance: sequences 1..2 generated alone with row_offset=1 give the same tokens
    mg.row_offset = 1
    mg.transformer.encode_text = lambda texts: te[1:1 + len(texts)]
    _, idc = mg.generate(texts=["a"] * 2, timesteps=8, return_ids=True)
    assert torch.equal(idc, ida[1:])


def make_branch_maskgit(name, precision):
    """MaskGit variants of the optional-branch fixtures G6-G9 (tests/golden/make_golden.py)."""
    m = M()
    vae = small_vae(precision)
    tr_sc = lambda: load(m.MaskGitTransformer(num_tokens=1024, dim=128, seq_len=16, depth=2, heads=2, t5_name="synth-128",
                                              self_cond=True, precision=precision),
                         util.transformer_sd(1024, 128, 16, 2, 2, seed=21, text_dim=128))
    if name == "gen_selfcond_small":
        return m.MaskGit(image_size=16, transformer=tr_sc(), vae=vae).cuda(), {}, 779
    if name in ("gen_critic_small", "gen_critic_forced_off_small"):
        critic = load(m.TokenCritic(num_tokens=1024, dim=128, seq_len=16, depth=1, heads=2, t5_name="synth-128", precision=precision),
                      util.critic_sd(1024, 128, 16, 1, 2, seed=22, text_dim=128))
        mg = m.MaskGit(image_size=16, transformer=small_transformer(precision), vae=vae, token_critic=critic).cuda()
        kw = dict(critic_noise_scale=0.7) if name == "gen_critic_small" else dict(force_not_use_token_critic=True)
        return mg, kw, 780
    if name == "gen_remask_prev_small":
        mg = m.MaskGit(image_size=16, transformer=small_transformer(precision), vae=vae, no_mask_token_prob=0.1).cuda()
        return mg, dict(can_remask_prev_masked=True), 782
    mg = m.MaskGit(image_size=16, transformer=tr_sc(), vae=vae, self_token_critic=True)
    w, b = util.self_critic_head()
    mg.token_critic.to_pred.weight.data.copy_(w); mg.token_critic.to_pred.bias.data.copy_(b)
    assert "token_critic.to_pred.weight" in mg.state_dict() and "token_critic.net.to_logits.weight" in mg.state_dict()
    return mg.cuda(), {}, 781


BRANCHES = ["gen_selfcond_small", "gen_critic_small", "gen_critic_forced_off_small", "gen_selfcritic_small", "gen_remask_prev_small"]


@pytest.mark.parametrize("name", BRANCHES)
def test_generate_optional_branches_fp32_token_identical_to_reference(name):
    """self-conditioning feedback, TokenCritic / SelfCritic scoring and can_remask_prev_masked: token ids
    bit-identical to the unmodified reference with its noise stream injected (gumbel draw, then critic draw, per step)."""
    g = util.golden(name)
    mg, kw, seed = make_branch_maskgit(name, "fp32")
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    mg.transformer.encode_text = lambda texts: te
    mg.sampler_noise_fn = util.torch_noise_fn(seed)
    images, ids = mg.generate(texts=["a"] * 3, timesteps=8, return_ids=True, **kw)
    assert torch.equal(ids.cpu(), g["ids"]), (ids.cpu() != g["ids"]).sum()
    assert (images.cpu() - g["images"]).abs().max() < 1e-4


@pytest.mark.parametrize("name", ["gen_selfcond_small", "gen_critic_small", "gen_selfcritic_small", "gen_remask_prev_small"])
def test_generate_optional_branches_bf16_graph_equals_eager(name):
    """bf16 + in-kernel Philox: the CUDA-graph replay and the eager launch sequence produce the same tokens, and a batch shard
    generated alone (row_offset) reproduces its rows (the critic noise is keyed by global position too)."""
    mg, kw, _ = make_branch_maskgit(name, "bf16")
    te = util.text_embeds("g4.te", 3, 8, 128, 14).cuda()
    mg.transformer.encode_text = lambda texts: te[:len(texts)]
    mg.sampler_seed = 5
    a, ida = mg.generate(texts=["a"] * 3, timesteps=8, return_ids=True, **kw)
    a2, ida2 = mg.generate(texts=["a"] * 3, timesteps=8, return_ids=True, **kw)        # replay
    mg.use_cuda_graph = False
    b_, idb = mg.generate(texts=["a"] * 3, timesteps=8, return_ids=True, **kw)
    assert torch.equal(ida, idb) and torch.equal(ida, ida2) and torch.equal(a, b_) and torch.isfinite(a).all()
    assert int(ida.min()) >= 0 and int(ida.max()) < 1024
    mg.row_offset = 1
    mg.transformer.encode_text = lambda texts: te[1:1 + len(texts)]
    _, idc = mg.generate(texts=["a"] * 2, timesteps=8, return_ids=True, **kw)
    assert torch.equal(idc, ida[1:])


def test_token_critic_forward_vs_oracle():
    """TokenCritic.forward_with_cond_scale (dim_out = 1 head) at the class boundary."""
    m = M()
    sd = util.critic_sd(1024, 128, 16, 1, 2, seed=22, text_dim=128)
    critic = load(m.TokenCritic(num_tokens=1024, dim=128, seq_len=16, depth=1, heads=2, t5_name="synth-128", precision="fp32"), sd)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    ids = torch.from_numpy((synth.uniform("crit.ids", (3, 16), 3) * 1024).astype(np.int64))
    want, _ = O.forward_with_cond_scale(sd, dict(heads=2, depth=1), ids, te, None, 3.0)
    got = critic.forward_with_cond_scale(ids.cuda(), text_embeds=te.cuda(), cond_scale=3.)
    assert got.shape == (3, 16, 1) and (got.cpu() - want).abs().max() < 2e-4


def test_muse_cascade_runs_on_device():
    """Muse(base, superres): base generate -> superres generate conditioned on the low-res images (kept on the device)."""
    m = M()
    base = make_maskgit("bf16")
    sup = make_maskgit("bf16", superres=True)
    te = util.text_embeds("g4.te", 3, 8, 128, 14).cuda()
    base.transformer.encode_text = lambda texts: te[:len(texts)]
    sup.transformer.encode_text = lambda texts: te[:len(texts)]
    base.sampler_seed, sup.sampler_seed = 7, 8
    muse = m.Muse(base=base, superres=sup)
    hi, lo = muse(["a", "b"], timesteps=6, return_lowres=True, return_pil_images=False)
    assert lo.shape == (2, 3, 16, 16) and hi.shape == (2, 3, 32, 32) and torch.isfinite(hi).all()
    pil = muse(["a", "b"], timesteps=4, return_pil_images=True)
    assert len(pil) == 2 and pil[0].size == (32, 32)
    # the cascade equals the two generate() calls made by hand
    lo2 = base.generate(["a", "b"], timesteps=6, cond_scale=3., temperature=1.)
    hi2 = sup.generate(["a", "b"], cond_images=lo2, timesteps=6, cond_scale=3., temperature=1.)
    assert torch.equal(lo, lo2) and torch.equal(hi, hi2)


@pytest.mark.parametrize("superres", [False, True], ids=["base", "superres"])
def test_decode_step_c_abi_equals_python_orchestration(superres):
    """mmg_decode_step (the whole decode step issued from C++: one call per step) launches exactly what the Python mirror launches
    call by call: same tokens and pixels, eager and under the whole-call CUDA graph, with injected noise and with Philox."""
    mg = make_maskgit("bf16", superres=superres)
    nb_ = 2 if superres else 3
    te = util.text_embeds("g4.te", 3, 8, 128, 14).cuda()[:nb_]
    mg.transformer.encode_text = lambda texts: te[:len(texts)]
    kw = dict(timesteps=8)
    if superres:
        kw["cond_images"] = torch.from_numpy(synth.uniform("g5.cond", (2, 3, 16, 16), 15)).cuda()
    outs = {}
    for native in (False, True):
        mg.use_native_step = native
        for graph in (False, True):
            mg.use_cuda_graph = graph
            mg.sampler_seed, mg.sampler_noise_fn = 9, None
            outs[(native, graph)] = mg.generate(texts=["a"] * nb_, return_ids=True, **kw)
        mg.use_cuda_graph = False
        mg.sampler_noise_fn = util.torch_noise_fn(778)
        outs[(native, "inject")] = mg.generate(texts=["a"] * nb_, return_ids=True, **kw)
        mg.sampler_noise_fn = None
    for key in (False, True, "inject"):
        (ia, ta), (ib, tb) = outs[(False, key)], outs[(True, key)]
        assert torch.equal(ta, tb), (key, int((ta != tb).sum()))
        assert torch.equal(ia, ib), key
    assert torch.equal(outs[(True, False)][1], outs[(True, True)][1])


def test_ff_geglu_c_abi_vs_torch():
    """mmg_ff_geglu: x += FeedForward(x) (LayerNorm -> GEGLU -> LayerNorm folded through the second product), optional constant added to
    the tail rows first (the null-CFG cross-attention term)."""
    from muse_maskgit_pytorch_b200 import ops
    import torch.nn.functional as Fn
    torch.manual_seed(5)
    R, dim, F_ = 300, 128, 341
    Fp = 384
    bf = torch.bfloat16
    x = torch.randn(R, dim) * 1.5
    g0, g3 = 1 + 0.1 * torch.randn(dim), 1 + 0.1 * torch.randn(F_)
    w1 = (torch.randn(2 * F_, dim) * dim ** -0.5).to(bf).float()
    w2 = (torch.randn(dim, F_) * F_ ** -0.5)
    add = torch.randn(dim) * 0.3
    split = 200
    # reference (fp32 math on the bf16-rounded operands the kernels see)
    xr = x.clone(); xr[split:] += add
    xn = Fn.layer_norm(xr, (dim,), g0, None, 1e-5).to(bf).float()
    hcat = xn @ w1.t()
    hh = (hcat[:, F_:] * Fn.gelu(hcat[:, :F_])).to(bf).float()
    w2f = torch.zeros(dim, Fp); w2f[:, :F_] = w2 * g3[None]
    w2f_b = w2f.to(bf)
    want = xr + Fn.layer_norm(hh, (F_,), None, None, 1e-5) @ w2f_b.float()[:, :F_].t()
    # packed operands (W1 rows interleaved in blocks of 32: [x(32) | gate(32)])
    w1p = torch.zeros(2 * Fp, dim)
    u = torch.arange(Fp); dst = u // 32 * 64 + u % 32; ok = u < F_
    w1p[dst[ok]] = w1[:F_]; w1p[dst[ok] + 32] = w1[F_:]
    xd = x.cuda()
    ops.ff_geglu(xd, g0.cuda(), w1p.to(bf).cuda(), w2f_b.cuda(), w2f_b.float().sum(1).cuda(), F_,
                 torch.empty(R, dim, dtype=bf, device="cuda"), torch.empty(R, Fp, dtype=bf, device="cuda"), torch.empty(R, Fp // 32, 2, device="cuda"),
                 add=add.cuda(), add_from=split)
    err = (xd.cpu() - want).abs().max().item()
    assert err < 3e-2, err
    assert rel_l2(xd, want) < 5e-3


def test_pack_cache_round_trip(tmp_path):
    """On-disk pre-pack cache: the first model writes its packed weights, a second instance with the same checkpoint reads
    them back (its own packing code is made to fail) and generates the same tokens; changed weights miss the cache."""
    from muse_maskgit_pytorch_b200 import pack_cache
    m = M()
    pack_cache.set_pack_cache(tmp_path)
    try:
        before = dict(pack_cache.stats)
        mg = make_maskgit("bf16")
        te = util.text_embeds("g4.te", 3, 8, 128, 14).cuda()
        mg.transformer.encode_text = lambda texts: te[:len(texts)]
        mg.sampler_seed = 5
        img_a, ids_a = mg.generate(["a"] * 3, timesteps=6, return_ids=True)
        assert pack_cache.stats["stores"] >= before["stores"] + 2 and len(list(tmp_path.iterdir())) >= 2      # transformer + VAE
        mg2 = make_maskgit("bf16")
        boom = lambda *a, **k: (_ for _ in ()).throw(AssertionError("packing ran although the cache holds this checkpoint"))
        mg2.transformer._build_pack = boom
        mg2.vae._build_pack = boom
        mg2.transformer.encode_text = lambda texts: te[:len(texts)]
        mg2.sampler_seed = 5
        img_b, ids_b = mg2.generate(["a"] * 3, timesteps=6, return_ids=True)
        assert torch.equal(ids_a, ids_b) and torch.equal(img_a, img_b)
        assert pack_cache.stats["hits"] >= before["hits"] + 2
        with torch.no_grad():
            mg2.transformer.to_logits.weight.mul_(1.5)
        misses = pack_cache.stats["misses"]
        del mg2.transformer._build_pack
        mg2.generate(["a"] * 3, timesteps=2)
        assert pack_cache.stats["misses"] == misses + 1
    finally:
        pack_cache.set_pack_cache(None)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vae_non_default_layout_vs_reference_golden(precision):
    """Encoder / decoder layouts other than the default (vqgan_vae.py:185-232): layer_mults (2, 4), per-stage res-block counts (1, 2) and a 3x3 stem; the fixture
    (made by the unmodified reference) carries its own weights, which also checks the state_dict keys of the generalised layout."""
    g = util.golden("vae_variant")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    vae = M().VQGanVAE(dim=16, layers=2, codebook_size=256, encdec_layer_mults=(2, 4), encdec_num_resnet_blocks=(1, 2), encdec_first_conv_kernel_size=3,
                       precision=precision)
    assert set(vae.state_dict()) == set(sd), set(vae.state_dict()) ^ set(sd)
    vae.load_state_dict(sd)
    vae = vae.cuda()
    ids = vae.encode_ids(g["img"].cuda())
    rec = vae.decode_from_ids(g["ids"].cuda())
    if precision == "fp32":
        assert torch.equal(ids.cpu(), g["ids"])
        assert (rec.cpu() - g["recon"]).abs().max() < 1e-4
    else:
        assert (ids.cpu() == g["ids"]).float().mean() > 0.9
        assert (rec.cpu() - g["recon"]).abs().max() < 5e-2


def test_generate_accepts_any_topk_threshold():
    """generate(topk_filter_thres=0.5) (k = 512 of 1024 here; at V = 65536 the same call keeps 32768 logits per row): runs and equals the oracle in
    parity precision."""
    mg = make_maskgit("fp32")
    sd = util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128)
    vsd = util.vae_sd(16, 2, 1024, seed=12)
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    mg.transformer.encode_text = lambda texts: te
    mg.sampler_noise_fn = util.torch_noise_fn(779)
    images, ids = mg.generate(texts=["a"] * 3, timesteps=6, topk_filter_thres=0.5, return_ids=True)
    ref_img, ref_ids = O.generate(sd, CFG, vsd, 10, te, 4, util.torch_noise_fn(779), timesteps=6, topk_thres=0.5)
    assert torch.equal(ids.cpu().view(3, -1), ref_ids)
    assert (images.cpu() - ref_img).abs().max() < 1e-4
