"""TEST INFRASTRUCTURE (imported only by tests/): the selection logic of libmmg's fused logits + sampling tail restated in numpy, to show on
the CPU that it returns exactly what the reference's  argmax(top_k(logits) / T + gumbel(u))  returns (muse_maskgit_pytorch.py:403-418, 576-582).

The fused path (csrc/mmg_logits_fused.cu) never forms a row's V logits.  It keeps (a) every logit >= a per-row threshold t chosen from a
4096-column sample so that about k + 4 sigma logits pass (the CANDIDATES: a superset of the exact top-k whenever at least k pass) and (b) the
row's softmax statistics.  The finisher then
  pass 1: perturbs every candidate that has not been excluded, p_v = x_v / T + g_v (g_v a pure function of (row, v)), takes the argmax
          (ties -> lowest vocabulary index);
  pass 2: counts the candidates that beat the winner's LOGIT (x > x_w, or x == x_w with a lower index) = its exact rank in the whole row,
          because every logit above x_w is >= t and therefore in the list;
  rank < k -> done; else the winner lies between t and the true k-th value: exclude it and repeat pass 1.
torch.topk's choice among logits equal to the k-th value is unspecified; libmmg keeps the lowest vocabulary indices."""
import numpy as np


def reference_choice(logits, gumbel, k, temperature):
    """argmax over the top-k set (k largest logits, ties at the k-th value -> lowest indices) of logits / T + gumbel; ties -> lowest index."""
    order = np.lexsort((np.arange(logits.size), -logits.astype(np.float64)))          # by logit descending, then index ascending
    keep = order[:k]
    p = logits[keep].astype(np.float32) / np.float32(max(temperature, 1e-10)) + gumbel[keep].astype(np.float32)
    best = np.lexsort((keep, -p.astype(np.float64)))[0]
    return int(keep[best])


def fused_choice(logits, gumbel, k, temperature, threshold, max_excl=24):
    """The finisher's procedure on the candidate list {v : logits[v] >= threshold}.  Returns (token, passes) or (None, passes) when the list
    is too short (the kernel then sends the row through the materialised path)."""
    idx = np.nonzero(logits >= threshold)[0]
    if idx.size < k:
        return None, 0
    x = logits[idx].astype(np.float32)
    p = x / np.float32(max(temperature, 1e-10)) + gumbel[idx].astype(np.float32)
    excluded = np.zeros(idx.size, dtype=bool)
    for n_pass in range(1, max_excl + 2):
        live = np.nonzero(~excluded)[0]
        w = live[np.lexsort((idx[live], -p[live].astype(np.float64)))[0]]               # pass 1
        rank = int(np.sum((x > x[w]) | ((x == x[w]) & (idx < idx[w]))))                 # pass 2
        if rank < k:
            return int(idx[w]), n_pass
        excluded[w] = True
    return None, max_excl + 1
