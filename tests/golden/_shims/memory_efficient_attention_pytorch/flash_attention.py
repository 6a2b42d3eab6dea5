"""Test-only stand-in for memory-efficient-attention-pytorch (>=0.1.4): FlashAttentionFunction.forward
restated from the published tiled online-softmax algorithm."""
import torch


class FlashAttentionFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, causal, q_bucket, k_bucket):
        scale = q.shape[-1] ** -0.5
        neg = -torch.finfo(q.dtype).max
        o = torch.zeros_like(q)
        n, j = q.shape[-2], k.shape[-2]
        for qs in range(0, n, q_bucket):
            qc = q[..., qs:qs + q_bucket, :]
            rows = qc.shape[-2]
            row_sum = torch.zeros((*q.shape[:-2], rows, 1), dtype=torch.float32)
            row_max = torch.full((*q.shape[:-2], rows, 1), neg, dtype=torch.float32)
            oc = torch.zeros_like(qc)
            for ks in range(0, j, k_bucket):
                kc, vc = k[..., ks:ks + k_bucket, :], v[..., ks:ks + k_bucket, :]
                s = torch.einsum("...id,...jd->...ij", qc, kc) * scale
                m = None
                if mask is not None:
                    m = mask[..., qs:qs + q_bucket, ks:ks + k_bucket] if mask.ndim == 4 else mask[:, None, None, ks:ks + k_bucket]
                    s = s.masked_fill(~m, neg)
                new_max = torch.maximum(s.amax(-1, keepdim=True), row_max)
                p = torch.exp(s - new_max)
                if m is not None:
                    p = p.masked_fill(~m, 0.)
                blk = p.sum(-1, keepdim=True).clamp(min=1e-10)
                corr = torch.exp(row_max - new_max)
                oc = oc * corr + torch.einsum("...ij,...jd->...id", p, vc)
                row_sum = corr * row_sum + blk
                row_max = new_max
            o[..., qs:qs + q_bucket, :] = oc / row_sum
        return o
