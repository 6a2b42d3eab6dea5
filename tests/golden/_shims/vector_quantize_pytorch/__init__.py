"""Test-only stand-in for vector-quantize-pytorch (>=1.11.8, not installed, no network): inference
branches of LFQ and (Euclidean) VectorQuantize restated from the published algorithm
(SURVEY.md Appendix A.1/A.2).  Only tests/golden/make_golden.py puts this on sys.path."""
import math
import torch
from torch import nn


class LFQ(nn.Module):
    def __init__(self, *, dim, codebook_size, diversity_gamma=1., **kw):
        super().__init__()
        d = int(math.log2(codebook_size))
        assert 2 ** d == codebook_size
        self.codebook_dim = d
        self.project_in = nn.Linear(dim, d) if dim != d else nn.Identity()
        self.project_out = nn.Linear(d, dim) if dim != d else nn.Identity()
        self.register_buffer("mask", 2 ** torch.arange(d - 1, -1, -1))

    def indices_to_codes(self, indices, project_out=True):
        bits = ((indices[..., None].int() & self.mask) != 0).float()
        codes = bits * 2 - 1
        return self.project_out(codes) if project_out else codes

    def forward(self, x):
        b, D, h, w = x.shape
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, D)
        x = self.project_in(x)
        pos = x > 0
        ids = (pos.int() * self.mask.int()).sum(-1)
        q = torch.where(pos, torch.ones_like(x), -torch.ones_like(x))
        out = self.project_out(q).reshape(b, h, w, D).permute(0, 3, 1, 2)
        return out, ids.reshape(b, h, w), torch.zeros((), device=x.device)


class VectorQuantize(nn.Module):
    def __init__(self, *, dim, codebook_size, accept_image_fmap=False, **kw):
        super().__init__()
        self.embed = nn.Parameter(torch.randn(codebook_size, dim))
        self.project_in = nn.Identity()
        self.project_out = nn.Identity()

    @property
    def codebook(self):
        return self.embed

    def forward(self, x):
        b, D, h, w = x.shape
        flat = x.permute(0, 2, 3, 1).reshape(-1, D)
        ids = (-torch.cdist(flat, self.embed)).argmax(-1)
        q = self.embed[ids].reshape(b, h, w, D).permute(0, 3, 1, 2)
        return q, ids.reshape(b, h, w), torch.zeros((), device=x.device)
