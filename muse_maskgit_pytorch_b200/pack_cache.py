"""On-disk cache of the packed (kernel-layout) weights, keyed by a hash of the module's parameters.

Packing = what `Transformer._packed()` / `VQGanVAE._packed()` derive from a state_dict: bf16 casts, the fused QKV matrix, GEGLU row
interleaving, LayerNorm folds, conv / conv-transpose tap layouts, the 3-way split of the LFQ projection.  It is redone whenever weights
change; with a cache directory set (`set_pack_cache(dir)` or MMG_PACK_CACHE=dir) a process that loads the same checkpoint again reads the
packed tensors back instead (`<dir>/<kind>-<blake2b of names, shapes, dtypes and bytes of every parameter and buffer>.pt`).
The cache holds derived data only: a miss, a truncated file or a version mismatch silently rebuilds."""
import hashlib
import os

import torch

PACK_FORMAT = 3          # bump when the packed layout of any kernel changes
_dir = os.environ.get("MMG_PACK_CACHE") or None
stats = {"hits": 0, "misses": 0, "stores": 0}


def set_pack_cache(directory):
    """Enable (path) or disable (None) the on-disk cache for this process."""
    global _dir
    _dir = os.fspath(directory) if directory else None


def weights_digest(module, extra=()):
    h = hashlib.blake2b(digest_size=20)
    h.update(repr((PACK_FORMAT, type(module).__name__) + tuple(extra)).encode())
    for name, t in sorted(list(module.named_parameters()) + list(module.named_buffers()), key=lambda kv: kv[0]):
        t = t.detach()
        h.update(repr((name, tuple(t.shape), str(t.dtype))).encode())
        h.update(t.contiguous().cpu().view(torch.uint8).numpy().tobytes() if t.numel() else b"")
    return h.hexdigest()


def _to(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: _to(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to(v, device) for v in obj)
    return obj


def load_or_build(kind, module, extra, build, device):
    """build() -> packed dict on `device`.  `extra`: everything besides the weights the packing depends on (precision, ...)."""
    if _dir is None:
        return build()
    path = os.path.join(_dir, f"{kind}-{weights_digest(module, extra)}.pt")
    if os.path.exists(path):
        try:
            packed = torch.load(path, map_location="cpu", weights_only=False)
            if packed.get("__format__") == PACK_FORMAT:
                stats["hits"] += 1
                packed.pop("__format__")
                return _to(packed, device)
        except Exception:
            pass                                   # unreadable / stale file: rebuild below and overwrite
    stats["misses"] += 1
    packed = build()
    try:
        os.makedirs(_dir, exist_ok=True)
        blob = _to({k: v for k, v in packed.items() if k != "sig"}, "cpu")
        blob["__format__"] = PACK_FORMAT
        tmp = path + f".tmp{os.getpid()}"
        torch.save(blob, tmp)
        os.replace(tmp, path)                      # atomic: concurrent ranks may race to write the same file
        stats["stores"] += 1
    except OSError:
        pass
    return packed
