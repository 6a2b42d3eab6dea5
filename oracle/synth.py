"""Deterministic synthetic tensors (TEST INFRASTRUCTURE — never imported by the product path).

Weights and inputs for the parity tests, the golden fixtures and the benchmark are generated
from a counter-based hash (splitmix64) keyed on the tensor *name*, so that

  * the golden-vector script (run in the build container, where /root/reference exists),
  * the CPU oracle, and
  * the GPU tests / bench (run on a box where /root/reference does not exist)

all see bit-identical fp32 values without shipping weight files.  Pure numpy, independent of the
torch RNG implementation and of module construction order.
"""
import zlib
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _key(name, seed):
    return np.uint64((zlib.crc32(name.encode()) << 32) ^ (seed & 0xFFFFFFFF))


def uniform(name, shape, seed=0):
    """U[0,1) float32 (24-bit mantissa), keyed on (name, seed)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + (_key(name, seed) << np.uint64(0))
        bits = _splitmix64(_splitmix64(ctr) ^ _key(name, seed))
    u = (bits >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))
    return u.reshape(shape)


def normal(name, shape, seed=0, std=1.0):
    """N(0, std^2) float32 via Box-Muller over two hashed uniforms (computed in float64, rounded once)."""
    u1 = uniform(name + "#1", shape, seed).astype(np.float64)
    u2 = uniform(name + "#2", shape, seed).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(1.0 - u1))
    return (r * np.cos(2.0 * np.pi * u2) * std).astype(np.float32)


def dyadic(name, shape, seed=0, bits=4, span=1.0):
    """Multiples of span/2^bits in [-span, span): products/sums of these stay exact in fp32 for the
    reduction lengths used in the bit-exact tests (any summation order gives the same bits)."""
    u = uniform(name, shape, seed)
    q = np.floor(u * (2 << bits)) - (1 << bits)
    return (q * (span / (1 << bits))).astype(np.float32)


def fill_state_dict(shapes, seed=0, overrides=None):
    """shapes: {param_name: shape}.  Returns {name: np.float32 array} with a name-dependent init:
    norm gains ~1, biases small, weights N(0, 1/fan_in)-ish so activations stay O(1)."""
    out = {}
    overrides = overrides or {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        if name in overrides:
            out[name] = overrides[name](name, shape)
            continue
        leaf = name.split(".")[-1]
        if leaf == "mask":  # LFQ bit weights 2^(d-1..0), integer buffer handled by caller
            continue
        if leaf in ("gamma", "q_scale", "k_scale") or (leaf == "weight" and len(shape) == 1):
            out[name] = (1.0 + 0.1 * normal(name, shape, seed)).astype(np.float32)
        elif leaf in ("beta",):
            out[name] = np.zeros(shape, np.float32)
        elif leaf == "bias":
            out[name] = (0.05 * normal(name, shape, seed)).astype(np.float32)
        elif leaf == "null_kv":
            out[name] = normal(name, shape, seed)
        elif "emb" in name:
            out[name] = normal(name, shape, seed)
        else:
            if len(shape) == 4 and "decoders" in name and ".0.weight" in name and "net" not in name \
                    and not name.endswith("decoders.0.weight"):
                fan_in = shape[0] * shape[2] * shape[3] / 4.0  # conv-transpose: (Cin, Cout, 4, 4), stride 2
            else:
                fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            out[name] = normal(name, shape, seed, std=1.0 / np.sqrt(max(fan_in, 1)))
    return out
