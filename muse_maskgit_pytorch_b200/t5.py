"""Text-encoder dimension lookup (mirror of the reference's t5.py surface that the hot path touches).

The benchmarked path consumes PRE-COMPUTED T5 embeddings (hook: assign `transformer.encode_text`), so only
`get_encoded_dim` / `T5_CONFIGS` are needed to construct a Transformer (ref: t5.py:44-55, muse_maskgit_pytorch.py:229-233).
`t5_encode_text` defers to HuggingFace transformers when a local model is available; it is outside the hot path.
"""
DEFAULT_T5_NAME = "google/t5-v1_1-base"
MAX_LENGTH = 256

# name -> {"config": object with .d_model} ; pre-seed to stay offline, e.g. T5_CONFIGS["t5-small"] = {"d_model": 512}
T5_CONFIGS = {}

_KNOWN_DIMS = {"t5-small": 512, "t5-base": 768, "t5-large": 1024, "t5-3b": 1024, "t5-11b": 1024,
               "google/t5-v1_1-small": 512, "google/t5-v1_1-base": 768, "google/t5-v1_1-large": 1024,
               "google/t5-v1_1-xl": 2048, "google/t5-v1_1-xxl": 4096}


def get_encoded_dim(name):
    entry = T5_CONFIGS.get(name)
    if entry is not None:
        if "d_model" in entry:
            return entry["d_model"]
        if "config" in entry:
            return entry["config"].d_model
        if "model" in entry:
            return entry["model"].config.d_model
    if name in _KNOWN_DIMS:
        return _KNOWN_DIMS[name]
    from transformers import T5Config          # needs a local HF cache; same behaviour as the reference
    cfg = T5Config.from_pretrained(name)
    T5_CONFIGS[name] = dict(config=cfg)
    return cfg.d_model


def t5_encode_text(texts, name=DEFAULT_T5_NAME, output_device=None):
    """ref: t5.py:60-99 — tokenizer + encoder, padded positions zero-filled.  Not part of the accelerated path."""
    import torch
    from transformers import T5Tokenizer, T5EncoderModel
    entry = T5_CONFIGS.setdefault(name, {})
    if "model" not in entry:
        entry["model"] = T5EncoderModel.from_pretrained(name)
        entry["tokenizer"] = T5Tokenizer.from_pretrained(name)
    model, tok = entry["model"], entry["tokenizer"]
    if torch.cuda.is_available():
        model = model.cuda()
    dev = next(model.parameters()).device
    enc = tok.batch_encode_plus(texts, return_tensors="pt", padding="longest", max_length=MAX_LENGTH, truncation=True)
    ids, attn = enc.input_ids.to(dev), enc.attention_mask.to(dev)
    model.eval()
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=attn).last_hidden_state.detach()
    out = out.masked_fill(~attn.bool()[..., None], 0.)
    return out if output_device is None else out.to(output_device)
