"""ATen-compatible sampler noise: the in-kernel stream equals what torch's CUDA generator draws for the
reference's `zeros_like(t).uniform_(0, 1)` calls, so `torch.manual_seed(s); generate()` reproduces a seeded GPU run of the
reference without materialising its [b, n, V] noise tensor."""
import numpy as np
import pytest
import torch

from oracle import philox, muse_oracle as O
from tests import util
from tests.test_gpu_models import make_maskgit, make_branch_maskgit

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _geometry():
    p = torch.cuda.get_device_properties(0)
    return p.multi_processor_count, p.max_threads_per_multi_processor


@pytest.mark.parametrize("numel", [3, 1000, 70001, 5 * 1184 * 256 + 17])
def test_model_of_aten_uniform_matches_torch_cuda(numel):
    """oracle/philox.aten_uniform (restated from the published ATen / cuRAND sources) against torch.cuda itself, at a non-zero
    generator offset, across the single-round, partial-grid and multi-round launch geometries."""
    sms, mt = _geometry()
    torch.manual_seed(1234)
    gen = torch.cuda.default_generators[0]
    first = torch.empty(numel, device="cuda").uniform_(0, 1).cpu().numpy()
    off1 = gen.get_offset()
    second = torch.empty(numel, device="cuda").uniform_(0, 1).cpu().numpy()
    stride = philox.aten_stride(numel, sms, mt)
    assert off1 == philox.aten_offset_increment(numel, stride), (off1, stride)
    assert gen.get_offset() == 2 * off1
    assert np.array_equal(first, philox.aten_uniform(1234, 0, numel, stride))
    assert np.array_equal(second, philox.aten_uniform(1234, off1, numel, stride))


def cuda_noise_fn(seed):
    """The draws a seeded GPU run of the reference makes (in call order), brought to the CPU oracle."""
    state = {"started": False}

    def fn(step, shape):
        if not state["started"]:
            torch.manual_seed(seed)
            state["started"] = True
        return torch.zeros(shape, device="cuda").uniform_(0, 1).cpu()
    return fn


def test_logits_sample_aten_mode_equals_injected_torch_noise():
    from muse_maskgit_pytorch_b200 import ops
    sms, mt = _geometry()
    B, n, V, nm, k = 3, 16, 4096, 5, 410
    torch.manual_seed(99)
    gen = torch.cuda.default_generators[0]
    torch.empty(8, device="cuda").uniform_(0, 1)                       # move the offset off zero
    off = gen.get_offset()
    u = torch.zeros((B, n, V), device="cuda").uniform_(0, 1)
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn((B * nm, V), generator=g) * 2).cuda()
    mp = torch.stack([torch.randperm(n, generator=g)[:nm].sort().values for _ in range(B)]).int().cuda().contiguous()
    outs = []
    for mode in ("inject", "aten"):
        ids = torch.full((B, n), V, dtype=torch.long, device="cuda")
        sc = torch.zeros((B, n), device="cuda")
        if mode == "inject":
            ops.logits_sample(logits, mp, ids, sc, nm, k, 0.7, u=u)
        else:
            seed_dev = torch.tensor([99], dtype=torch.int64, device="cuda")
            off_dev = torch.tensor([off], dtype=torch.int64, device="cuda")
            ops.logits_sample(logits, mp, ids, sc, nm, k, 0.7, seed=0, seed_dev=seed_dev,
                              aten=(0, off_dev, philox.aten_stride(B * n * V, sms, mt)))
        outs.append((ids.cpu(), sc.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("name", [None, "gen_critic_small", "gen_selfcritic_small"])
def test_generate_aten_rng_reproduces_seeded_gpu_reference_run(name):
    """fp32 generate() with sampler_rng="aten" under torch.manual_seed(s) == the CPU oracle fed with the tensors torch.cuda draws
    under the same seed (gumbel draw, then critic draw, per step); the CUDA generator ends where the reference would leave it."""
    te = util.text_embeds("g4.te", 3, 8, 128, 14)
    if name is None:
        mg, kw = make_maskgit("fp32"), {}
        sd, cfg, okw = util.transformer_sd(1024, 128, 16, 2, 2, seed=13, text_dim=128), dict(heads=2, depth=2), {}
    else:
        mg, kw, _ = make_branch_maskgit(name, "fp32")
        from tests.test_oracle_golden import _branch_case
        sd, cfg, okw, _ = _branch_case(name)
    want = O.generate_ids(sd, cfg, te, 16, 1024, cuda_noise_fn(4242), timesteps=8, **okw)
    end_offset = torch.cuda.default_generators[0].get_offset()
    mg.transformer.encode_text = lambda texts: te
    mg.sampler_rng = "aten"
    for graph in (True, False):
        mg.use_cuda_graph = graph
        torch.manual_seed(4242)
        _, ids = mg.generate(texts=["a"] * 3, timesteps=8, return_ids=True, **kw)
        assert torch.equal(ids.cpu().view(3, 16), want), (graph, (ids.cpu().view(3, 16) != want).sum())
        assert torch.cuda.default_generators[0].get_offset() == end_offset
