"""CPU ORACLE for the MaskGit.generate() hot path — TEST INFRASTRUCTURE ONLY.

A functional (state-dict in, tensor out) fp32 restatement of the reference algorithm on torch CPU.
Nothing under muse_maskgit_pytorch_b200/ may import this module; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do, and only as the checker / the CPU arm.

Pinning: tests/golden/make_golden.py runs the UNMODIFIED reference files from /root/reference (with
shims for its four missing third-party packages) on hash-generated weights and stores the outputs in
tests/golden/*.npz; tests/test_oracle_golden.py checks this oracle against those vectors.  The two
third-party functions on the path (vector-quantize-pytorch>=1.11.8 `LFQ`/`VectorQuantize`,
memory-efficient-attention-pytorch>=0.1.4 `FlashAttentionFunction`; lower-bound pins only, setup.py:25,32)
are not vendored in the reference, so at those two boundaries parity is pinned to the reference's own
non-flash `Attend` branch (attend.py:123-138) and to the LFQ == argmin-over-{+-1}^d identity.

All `ref:` citations are relative to /root/reference/muse_maskgit_pytorch/.
State-dict keys are exactly the reference modules' keys.
"""
import math
import torch
import torch.nn.functional as F

NEG_MAX = -torch.finfo(torch.float32).max


# ----------------------------------------------------------------------------- transformer blocks

def ln(x, gamma):
    """ref: muse_maskgit_pytorch.py:63-70 — F.layer_norm, eps 1e-5, learnable gamma, beta == 0 buffer."""
    return F.layer_norm(x, x.shape[-1:], gamma, torch.zeros_like(gamma))


def _heads(t, h):
    b, n, hd = t.shape
    return t.view(b, n, h, hd // h).transpose(1, 2)


def attention(sd, pre, x, heads, context=None, key_mask=None, scale=8.0):
    """ref: muse_maskgit_pytorch.py:126-162 (Attention.forward) + attend.py:123-138 (non-flash Attend).

    pre-LN on x only; k/v from `context` when cross-attending (context is NOT normed); the learned
    null key/value is prepended at key index 0 BEFORE the l2-norm; q,k are l2-normalised (eps 1e-12)
    then multiplied by q_scale/k_scale; logits = 8 * q.k; masked keys get -finfo.max; the null key is
    never masked (mask padded True on the left)."""
    xn = ln(x, sd[pre + "norm.gamma"])
    src = xn if context is None else context
    q = _heads(xn @ sd[pre + "to_q.weight"].t(), heads)
    k, v = (src @ sd[pre + "to_kv.weight"].t()).chunk(2, dim=-1)
    k, v = _heads(k, heads), _heads(v, heads)
    b = x.shape[0]
    null_k, null_v = sd[pre + "null_kv"][0], sd[pre + "null_kv"][1]          # (h, 1, dh)
    k = torch.cat((null_k.unsqueeze(0).expand(b, -1, -1, -1), k), dim=2)
    v = torch.cat((null_v.unsqueeze(0).expand(b, -1, -1, -1), v), dim=2)
    q = F.normalize(q, dim=-1) * sd[pre + "q_scale"]
    k = F.normalize(k, dim=-1) * sd[pre + "k_scale"]
    sim = (q @ k.transpose(-1, -2)) * scale
    if key_mask is not None:
        km = F.pad(key_mask, (1, 0), value=True)[:, None, None, :]
        sim = sim.masked_fill(~km, NEG_MAX)
    out = sim.softmax(dim=-1) @ v
    out = out.transpose(1, 2).reshape(b, x.shape[1], -1)
    return out @ sd[pre + "to_out.weight"].t()


def geglu_ff(sd, pre, x):
    """ref: muse_maskgit_pytorch.py:72-89 — LN -> Linear(d, 2F) -> gate * gelu_erf(x) -> LN(F) -> Linear(F, d).
    chunk(2): FIRST half goes through exact-erf GELU, SECOND half is the gate."""
    h = ln(x, sd[pre + "0.gamma"]) @ sd[pre + "1.weight"].t()
    a, gate = h.chunk(2, dim=-1)
    h = gate * F.gelu(a)
    return ln(h, sd[pre + "3.gamma"]) @ sd[pre + "4.weight"].t()


def transformer_blocks(sd, x, heads, depth, context, key_mask):
    """ref: muse_maskgit_pytorch.py:187-195."""
    for i in range(depth):
        p = f"transformer_blocks.layers.{i}."
        x = attention(sd, p + "0.", x, heads) + x
        x = attention(sd, p + "1.", x, heads, context=context, key_mask=key_mask) + x
        x = geglu_ff(sd, p + "2.", x) + x
    return ln(x, sd["transformer_blocks.norm.gamma"])


def transformer_forward(sd, cfg, ids, text_embeds, cond_ids=None, drop_text=False, self_cond_embed=None,
                        return_embed=False):
    """ref: muse_maskgit_pytorch.py:279-338 (inference branches).

    cfg: dict(heads, depth, self_cond).  drop_text=True is cond_drop_prob == 1 (prob_mask_like(.., 0.) ->
    all False, no RNG consumed: muse_maskgit_pytorch.py:308-310, 393-399).  `Transformer.norm` is NOT
    applied (defect B4) and to_logits has no bias."""
    b, n = ids.shape
    ctx = text_embeds @ sd["text_embed_proj.weight"].t() if "text_embed_proj.weight" in sd else text_embeds
    key_mask = (text_embeds != 0).any(dim=-1)
    if drop_text:
        key_mask = key_mask & torch.zeros(b, 1, dtype=torch.bool)
    if cond_ids is not None:
        cond_ids = cond_ids.reshape(b, -1)
        ctx = torch.cat((ctx, sd["token_emb.weight"][cond_ids]), dim=1)       # no pos-emb on cond tokens
        key_mask = F.pad(key_mask, (0, cond_ids.shape[-1]), value=True)
    x = sd["token_emb.weight"][ids] + sd["pos_emb.weight"][:n]
    if cfg.get("self_cond", False):
        sc = torch.zeros_like(x) if self_cond_embed is None else self_cond_embed
        x = x + geglu_ff(sd, "self_cond_to_init_embed.", sc)
    embed = transformer_blocks(sd, x, cfg["heads"], cfg["depth"], ctx, key_mask)
    logits = embed @ sd["to_logits.weight"].t()
    return (logits, embed) if return_embed else logits


def forward_with_cond_scale(sd, cfg, ids, text_embeds, cond_ids=None, cond_scale=3.0, self_cond_embed=None):
    """ref: muse_maskgit_pytorch.py:240-259.  Returns (scaled_logits, cond_embed)."""
    lc, embed = transformer_forward(sd, cfg, ids, text_embeds, cond_ids, False, self_cond_embed, True)
    if cond_scale == 1:
        return lc, embed
    ln_ = transformer_forward(sd, cfg, ids, text_embeds, cond_ids, True, self_cond_embed)
    return ln_ + (lc - ln_) * cond_scale, embed


# ----------------------------------------------------------------------------- sampler

def mask_schedule(seq_len, timesteps):
    """ref: muse_maskgit_pytorch.py:556-559, 422-423 — per step max(int(cos(t*pi/2)*seq), 1) with
    t = linspace(0,1,T) in fp32 (the last cos is ~ -4.4e-8 -> int 0 -> clamped to 1)."""
    t = torch.linspace(0, 1, timesteps)
    return [max(int((torch.cos(ti * math.pi * 0.5) * seq_len).item()), 1) for ti in t]


def top_k_count(vocab, thres):
    """ref: muse_maskgit_pytorch.py:414."""
    return math.ceil((1 - thres) * vocab)


def top_k_filter(logits, thres):
    """ref: muse_maskgit_pytorch.py:413-418."""
    k = top_k_count(logits.shape[-1], thres)
    val, ind = logits.topk(k, dim=-1)
    out = torch.full_like(logits, float("-inf"))
    out.scatter_(-1, ind, val)
    return out


def gumbel_from_uniform(u):
    """ref: muse_maskgit_pytorch.py:403-408 — -log(-log(u)) with log(x) = log(clamp(x, min=1e-20))."""
    lg = lambda t: torch.log(t.clamp(min=1e-20))
    return -lg(-lg(u))


def sample_step(logits, ids, mask_id, temperature, u, topk_thres=0.9, can_remask_prev_masked=False):
    """One sampling tail: ref muse_maskgit_pytorch.py:576-612 (no token critic).  can_remask_prev_masked=True keeps the
    confidence of already-decoded positions instead of pinning it to -1e5 (:609-612), so they compete for re-masking.
    `u` is the U[0,1) tensor the reference would draw with zeros_like(logits).uniform_(0,1).
    Returns (new_ids, new_scores, pred_ids)."""
    filtered = top_k_filter(logits, topk_thres)
    pred = (filtered / max(temperature, 1e-10) + gumbel_from_uniform(u)).argmax(dim=-1)
    is_mask = ids == mask_id
    new_ids = torch.where(is_mask, pred, ids)
    probs = logits.softmax(dim=-1)
    scores = 1 - probs.gather(2, pred[..., None])[..., 0]
    if not can_remask_prev_masked:
        scores = scores.masked_fill(~is_mask, -1e5)
    return new_ids, scores, pred


def remask(ids, scores, num_masked, mask_id):
    """ref: muse_maskgit_pytorch.py:561-563."""
    idx = scores.topk(num_masked, dim=-1).indices
    return ids.scatter(1, idx, mask_id)


def critic_scores(critic, ids, text_embeds, cond_ids, cond_scale, u, noise_mul):
    """Token-critic scoring branch, ref: muse_maskgit_pytorch.py:590-600.

    critic = dict(kind="token", sd, cfg)                 -> TokenCritic.forward_with_cond_scale (dim_out = 1 head, :383-386)
           | dict(kind="self", sd, cfg, w_pred, b_pred)  -> SelfCritic (:352-361): to_pred(embed of the COND forward); the CFG-combined
                                                            logits it also computes are discarded.
    No self_cond_embed is passed to the critic (zeros -> FeedForward(0) = 0).  u = the (b, n) uniform_ draw of `uniform(scores.shape)`."""
    if critic["kind"] == "token":
        logit, _ = forward_with_cond_scale(critic["sd"], critic["cfg"], ids, text_embeds, cond_ids, cond_scale)
        s = logit[..., 0]
    else:
        _, embed = transformer_forward(critic["sd"], critic["cfg"], ids, text_embeds, cond_ids, False, None, True)
        s = (embed @ critic["w_pred"].t() + critic["b_pred"])[..., 0]
    return s + (u - 0.5) * noise_mul


def generate_ids(sd, cfg, text_embeds, seq_len, mask_id, noise_fn, cond_ids=None, timesteps=18,
                 cond_scale=3.0, temperature=1.0, topk_thres=0.9, trace=None, max_steps=None,
                 critic=None, critic_noise_scale=1.0, can_remask_prev_masked=False):
    """ref: muse_maskgit_pytorch.py:507-613 (token loop of MaskGit.generate).

    noise_fn(step, shape) -> U[0,1) fp32 tensor, standing in for `zeros_like(t).uniform_(0, 1)`
    (muse_maskgit_pytorch.py:407) with shape (b, n, V), and — when a critic is given — for `uniform(scores.shape)`
    (:598) with shape (b, n), called in the reference's order (gumbel draw, then critic draw, every step).
    cfg["self_cond"]: the cond forward's embed is fed back through self_cond_to_init_embed on the next step (:574, 325-328).
    trace (list) receives per-step dicts for teacher-forced parity."""
    b = text_embeds.shape[0]
    ids = torch.full((b, seq_len), mask_id, dtype=torch.long)
    scores = torch.zeros((b, seq_len), dtype=torch.float32)
    sched = mask_schedule(seq_len, timesteps)
    self_cond_embed = None
    for step, (num_masked, steps_until_x0) in enumerate(zip(sched, reversed(range(timesteps)))):
        if max_steps is not None and step >= max_steps:      # bounded CPU-baseline sample (bench.py): stop early
            break
        ids = remask(ids, scores, num_masked, mask_id)
        logits, embed = forward_with_cond_scale(sd, cfg, ids, text_embeds, cond_ids, cond_scale, self_cond_embed)
        if cfg.get("self_cond", False):
            self_cond_embed = embed
        temp = temperature * (steps_until_x0 / timesteps)
        u = noise_fn(step, logits.shape)
        masked_in = ids
        ids, scores, pred = sample_step(logits, ids, mask_id, temp, u, topk_thres, can_remask_prev_masked)
        if critic is not None:
            uc = noise_fn(step, scores.shape)
            scores = critic_scores(critic, ids, text_embeds, cond_ids, cond_scale, uc,
                                   critic_noise_scale * (steps_until_x0 / timesteps))
        if trace is not None:
            trace.append(dict(ids_in=masked_in, logits=logits, u=u, temperature=temp, pred=pred,
                              ids_out=ids, scores=scores, num_masked=num_masked, embed=embed))
    return ids


# ----------------------------------------------------------------------------- VQGanVAE

def _lrelu(x):
    return F.leaky_relu(x, 0.1)          # ref: vqgan_vae.py:103-104 (slope hard-coded 0.1)


def vae_layers(sd):
    return sum(1 for k in sd if k.startswith("enc_dec.encoders.") and k.endswith(".0.weight")
               and k.count(".") == 4)


def vae_encode_fmap(sd, img, groups=16):
    """ref: vqgan_vae.py:241-244, 223-231, 267-281 — 5x5 conv, layers x (4x4 s2 conv + LeakyReLU), ResBlock."""
    L = vae_layers(sd)
    x = F.conv2d(img, sd["enc_dec.encoders.0.weight"], sd["enc_dec.encoders.0.bias"], padding=2)
    for i in range(1, L + 1):
        x = _lrelu(F.conv2d(x, sd[f"enc_dec.encoders.{i}.0.weight"], sd[f"enc_dec.encoders.{i}.0.bias"],
                            stride=2, padding=1))
    p = f"enc_dec.encoders.{L + 1}.net."
    if p + "0.weight" in sd:
        h = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1)
        h = _lrelu(F.group_norm(h, groups, sd[p + "1.weight"], sd[p + "1.bias"]))
        h = F.conv2d(h, sd[p + "3.weight"], sd[p + "3.bias"], padding=1)
        h = _lrelu(F.group_norm(h, groups, sd[p + "4.weight"], sd[p + "4.bias"]))
        x = F.conv2d(h, sd[p + "6.weight"], sd[p + "6.bias"]) + x
    return x


def vae_decode_fmap(sd, fmap, groups=16):
    """ref: vqgan_vae.py:246-265, 225, 232 — GLUResBlock, layers x (ConvTranspose 4x4 s2 p1 + LeakyReLU), 1x1 conv."""
    L = vae_layers(sd)
    x = fmap
    p = "enc_dec.decoders.0.net."
    if p + "0.weight" in sd:
        h = F.glu(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1), dim=1)
        h = F.group_norm(h, groups, sd[p + "2.weight"], sd[p + "2.bias"])
        h = F.glu(F.conv2d(h, sd[p + "3.weight"], sd[p + "3.bias"], padding=1), dim=1)
        h = F.group_norm(h, groups, sd[p + "5.weight"], sd[p + "5.bias"])
        x = F.conv2d(h, sd[p + "6.weight"], sd[p + "6.bias"]) + x
    for i in range(1, L + 1):
        x = _lrelu(F.conv_transpose2d(x, sd[f"enc_dec.decoders.{i}.0.weight"], sd[f"enc_dec.decoders.{i}.0.bias"],
                                      stride=2, padding=1))
    return F.conv2d(x, sd[f"enc_dec.decoders.{L + 1}.weight"], sd[f"enc_dec.decoders.{L + 1}.bias"])


def lfq_quantize(sd, fmap):
    """LFQ inference branch (vector-quantize-pytorch, restated; call site ref: vqgan_vae.py:331-335, 424).
    project_in -> sign -> ids = sum((x>0) * 2^(d-1-i)) (MSB first; x == 0 -> bit 0) -> +-1 -> project_out.
    Returns (quantized fmap (b,D,h,w), ids (b,h,w) int64)."""
    b, D, h, w = fmap.shape
    x = fmap.permute(0, 2, 3, 1).reshape(b, h * w, D)
    if "quantizer.project_in.weight" in sd:
        x = x @ sd["quantizer.project_in.weight"].t() + sd["quantizer.project_in.bias"]
    d = x.shape[-1]
    weights = 2 ** torch.arange(d - 1, -1, -1)
    pos = x > 0
    ids = (pos.long() * weights).sum(-1)
    q = torch.where(pos, 1.0, -1.0)
    if "quantizer.project_out.weight" in sd:
        q = q @ sd["quantizer.project_out.weight"].t() + sd["quantizer.project_out.bias"]
    return q.reshape(b, h, w, D).permute(0, 3, 1, 2), ids.reshape(b, h, w)


def lfq_codes_from_ids(sd, ids, code_bits):
    """LFQ.indices_to_codes(project_out=True) restated (call site ref: vqgan_vae.py:429-432). ids (..., ) -> (..., D)."""
    weights = 2 ** torch.arange(code_bits - 1, -1, -1)
    bits = (ids[..., None] & weights) != 0
    q = bits.float() * 2 - 1
    if "quantizer.project_out.weight" in sd:
        q = q @ sd["quantizer.project_out.weight"].t() + sd["quantizer.project_out.bias"]
    return q


def vq_l2_argmin(x, codebook):
    """Explicit Euclidean VQ lookup = the intended path behind the broken vqgan_vae.py:337-342 (defect B1):
    ids = argmin_k ||x - e_k||^2, first index on ties.  x (M, D), codebook (K, D)."""
    d2 = (x * x).sum(-1, keepdim=True) - 2 * x @ codebook.t() + (codebook * codebook).sum(-1)[None]
    return d2.argmin(dim=-1)


def vae_encode(sd, img):
    """ref: vqgan_vae.py:422-425 (LFQ default)."""
    fmap = vae_encode_fmap(sd, img)
    q, ids = lfq_quantize(sd, fmap)
    return q, ids


def vae_decode_from_ids(sd, ids, code_bits):
    """ref: vqgan_vae.py:427-438 (LFQ branch). ids (b,h,w) -> images (b,3,H,W), unclamped fp32."""
    codes = lfq_codes_from_ids(sd, ids, code_bits)            # (b,h,w,D)
    return vae_decode_fmap(sd, codes.permute(0, 3, 1, 2))


def generate(sd_tr, cfg, sd_vae, code_bits, text_embeds, fmap_size, noise_fn, cond_images=None, sd_cond_vae=None,
             **kw):
    """ref: muse_maskgit_pytorch.py:493-621 — full MaskGit.generate with pre-computed text embeddings."""
    cond_ids = None
    if cond_images is not None:
        _, cond_ids = vae_encode(sd_cond_vae if sd_cond_vae is not None else sd_vae, cond_images)
    V = sd_tr["to_logits.weight"].shape[0]
    ids = generate_ids(sd_tr, cfg, text_embeds, fmap_size * fmap_size, V, noise_fn, cond_ids=cond_ids, **kw)
    return vae_decode_from_ids(sd_vae, ids.view(-1, fmap_size, fmap_size), code_bits), ids
