"""State-dict key/shape tables of the reference modules (TEST INFRASTRUCTURE).

tests/golden/make_golden.py asserts these tables equal `module.state_dict()` of the unmodified reference
classes (ref: muse_maskgit_pytorch.py:199-238, vqgan_vae.py:185-232, 285-342), so the checkpoint-key
contract of SURVEY.md section 8(b) is pinned against the reference itself.
"""
import math


def ff_inner(dim, mult=4):
    return int(dim * mult * 2 / 3)          # ref: muse_maskgit_pytorch.py:82


def transformer_shapes(num_tokens, dim, seq_len, depth, heads=8, dim_head=64, ff_mult=4, text_dim=None,
                       add_mask_id=True, dim_out=None):
    inner = heads * dim_head
    F_ = ff_inner(dim, ff_mult)
    s = {"token_emb.weight": (num_tokens + int(add_mask_id), dim), "pos_emb.weight": (seq_len, dim)}
    for i in range(depth):
        for j in (0, 1):
            p = f"transformer_blocks.layers.{i}.{j}."
            s[p + "null_kv"] = (2, heads, 1, dim_head)
            s[p + "q_scale"] = (dim_head,)
            s[p + "k_scale"] = (dim_head,)
            s[p + "norm.gamma"] = (dim,)
            s[p + "norm.beta"] = (dim,)
            s[p + "to_q.weight"] = (inner, dim)
            s[p + "to_kv.weight"] = (2 * inner, dim)
            s[p + "to_out.weight"] = (dim, inner)
        p = f"transformer_blocks.layers.{i}.2."
        s[p + "0.gamma"] = (dim,); s[p + "0.beta"] = (dim,)
        s[p + "1.weight"] = (2 * F_, dim)
        s[p + "3.gamma"] = (F_,); s[p + "3.beta"] = (F_,)
        s[p + "4.weight"] = (dim, F_)
    s["transformer_blocks.norm.gamma"] = (dim,); s["transformer_blocks.norm.beta"] = (dim,)
    s["norm.gamma"] = (dim,); s["norm.beta"] = (dim,)
    s["to_logits.weight"] = (dim_out if dim_out is not None else num_tokens, dim)
    if text_dim is not None and text_dim != dim:
        s["text_embed_proj.weight"] = (dim, text_dim)
    F0 = ff_inner(dim, 4)
    p = "self_cond_to_init_embed."
    s[p + "0.gamma"] = (dim,); s[p + "0.beta"] = (dim,); s[p + "1.weight"] = (2 * F0, dim)
    s[p + "3.gamma"] = (F0,); s[p + "3.beta"] = (F0,); s[p + "4.weight"] = (dim, F0)
    return s


def vae_shapes(dim, channels=3, layers=4, codebook_size=65536, lfq=True, resblocks=1):
    dims = [dim] + [dim * 2 ** i for i in range(layers)]
    D = dims[-1]
    s = {"enc_dec.encoders.0.weight": (dim, channels, 5, 5), "enc_dec.encoders.0.bias": (dim,)}
    for i in range(layers):
        s[f"enc_dec.encoders.{i + 1}.0.weight"] = (dims[i + 1], dims[i], 4, 4)
        s[f"enc_dec.encoders.{i + 1}.0.bias"] = (dims[i + 1],)
    if resblocks:
        p = f"enc_dec.encoders.{layers + 1}.net."
        for j in (0, 3):
            s[p + f"{j}.weight"] = (D, D, 3, 3); s[p + f"{j}.bias"] = (D,)
        for j in (1, 4):
            s[p + f"{j}.weight"] = (D,); s[p + f"{j}.bias"] = (D,)
        s[p + "6.weight"] = (D, D, 1, 1); s[p + "6.bias"] = (D,)
        p = "enc_dec.decoders.0.net."
        for j in (0, 3):
            s[p + f"{j}.weight"] = (2 * D, D, 3, 3); s[p + f"{j}.bias"] = (2 * D,)
        for j in (2, 5):
            s[p + f"{j}.weight"] = (D,); s[p + f"{j}.bias"] = (D,)
        s[p + "6.weight"] = (D, D, 1, 1); s[p + "6.bias"] = (D,)
    for i in range(layers):                      # decoders are prepended: decoders.1 undoes the LAST encoder stage
        cin, cout = dims[layers - i], dims[layers - i - 1]
        s[f"enc_dec.decoders.{i + 1}.0.weight"] = (cin, cout, 4, 4)      # ConvTranspose2d layout (Cin, Cout, kh, kw)
        s[f"enc_dec.decoders.{i + 1}.0.bias"] = (cout,)
    s[f"enc_dec.decoders.{layers + 1}.weight"] = (channels, dim, 1, 1)
    s[f"enc_dec.decoders.{layers + 1}.bias"] = (channels,)
    if lfq:
        d = int(math.log2(codebook_size))
        assert 2 ** d == codebook_size
        if d != D:
            s["quantizer.project_in.weight"] = (d, D); s["quantizer.project_in.bias"] = (d,)
            s["quantizer.project_out.weight"] = (D, d); s["quantizer.project_out.bias"] = (D,)
    return s
