"""Test-scaffolding stand-in for `torch_scatter` (absent from this image).

ORACLE INFRASTRUCTURE ONLY -- never imported by the product path.
Implements the three entry points the reference's PaiNN path calls
(painn_denoising.py:38,565-566,299; utils/utils.py:762,773;
denoising_torch.py:461) with plain torch ops that are mathematically the
same reduction (real torch_scatter's `scatter_sum` is `Tensor.scatter_add_`).
"""
import torch


def _expand_index(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if dim < 0:
        dim = src.dim() + dim
    idx = _expand_index(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        res = torch.zeros(size, dtype=src.dtype, device=src.device)
        return res.scatter_add_(dim, idx, src)
    if reduce == "mean":
        res = torch.zeros(size, dtype=src.dtype, device=src.device)
        res.scatter_add_(dim, idx, src)
        ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt.scatter_add_(0, index, ones)
        cnt.clamp_(min=1)
        shape = [1] * src.dim()
        shape[dim] = dim_size
        cnt = cnt.view(shape)
        if res.is_floating_point():
            return res.true_divide(cnt)
        return res.div(cnt, rounding_mode="floor")
    if reduce in ("min", "max"):
        res = torch.zeros(size, dtype=src.dtype, device=src.device)
        return res.scatter_reduce_(
            dim, idx, src, reduce="a" + reduce, include_self=False
        )
    raise ValueError(reduce)


def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    assert reduce == "sum"
    return scatter(src, index, dim=0, dim_size=dim_size, reduce="sum")


def segment_csr(src, indptr, out=None, reduce="sum"):
    assert reduce == "sum"
    csum = torch.cat([src.new_zeros(1), src.cumsum(0)])
    return csum[indptr[1:]] - csum[indptr[:-1]]
