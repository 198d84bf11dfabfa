"""Symmetric (generalised) eigensolver with a broadened backward pass (the published TBMaLT ``eighb``: degenerate
eigenvalue pairs get a finite 1/(e_j - e_i) in the eigenvector gradient) and identity padding for zero-padded batches."""
import torch

__all__ = ["eighb"]


class _SymEigB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, method="cond", factor=1e-12):
        if torch.is_tensor(factor):
            factor = float(factor)
        w, v = torch.linalg.eigh(a)
        ctx.save_for_backward(w, v)
        ctx.bm, ctx.bf = method, factor
        return w, v

    @staticmethod
    def backward(ctx, w_bar, v_bar):
        w, v = ctx.saved_tensors
        bm, bf = ctx.bm, ctx.bf
        vt = v.transpose(-2, -1)
        deltas = w.unsqueeze(-2) - w.unsqueeze(-1)  # [i, j] = w_j - w_i
        if bm == "cond":
            big = torch.abs(deltas) > bf
            f = torch.where(big, 1.0 / torch.where(big, deltas, torch.ones_like(deltas)), torch.sign(deltas) / bf)
        elif bm == "lorn":
            f = deltas / (deltas**2 + bf)
        else:
            f = 1.0 / deltas
        eye = torch.eye(w.shape[-1], dtype=torch.bool, device=w.device)
        f = torch.where(eye, torch.zeros_like(f), f)
        inner = torch.diag_embed(w_bar) + f * (vt @ v_bar)
        a_bar = v @ inner @ vt
        a_bar = 0.5 * (a_bar + a_bar.transpose(-2, -1))
        return a_bar, None, None


def _eig_sort_out(w, v, ghost=True):
    """Move the padding ("ghost") eigenpairs to the end and zero their eigenvalues."""
    big = torch.finfo(w.dtype).max
    is_ghost = torch.eq(w, 0) if ghost else torch.eq(w, 1)
    count = is_ghost.sum(-1)
    w_ = torch.where(is_ghost, torch.full_like(w, big), w)
    order = torch.argsort(w_, dim=-1)
    w = torch.gather(w, -1, order)
    v = torch.gather(v, -1, order.unsqueeze(-2).expand_as(v))
    n = w.shape[-1]
    tail = torch.arange(n, device=w.device).expand_as(w) >= (n - count).unsqueeze(-1)
    w = torch.where(tail, torch.zeros_like(w), w)
    return w, v


def _eigh(a, method, factor):
    if method is None:  # no gradient through the eigenvectors requested
        return torch.linalg.eigh(a)
    return _SymEigB.apply(a, method, factor)


def eighb(a, b=None, scheme="chol", broadening_method="cond", factor=1e-12, sort_out=True, aux=True, is_posdef=False, **kwargs):
    if b is None:
        w, v = _eigh(a, broadening_method, factor)
    else:
        if aux:  # zero padding -> identity padding, so that B stays positive definite
            is_zero = torch.eq(b, 0)
            mask = torch.all(is_zero, dim=-1) & torch.all(is_zero, dim=-2)
            b = b + torch.diag_embed(mask.type(a.dtype))
        if scheme == "chol":
            l = torch.linalg.cholesky(b)
            if kwargs.get("direct_inv", False):
                l_inv = torch.inverse(l)
            else:
                eye = torch.eye(a.shape[-1], dtype=a.dtype, device=b.device)
                l_inv = torch.linalg.solve_triangular(l, eye.expand_as(l) if a.ndim > 2 else eye, upper=False)
            l_inv_t = torch.transpose(l_inv, -1, -2)
            c = l_inv @ a @ l_inv_t
            w, v_ = _eigh(c, broadening_method, factor)
            v = l_inv_t @ v_
        elif scheme == "lowd":
            wb, vb = torch.linalg.eigh(b)
            b_so = vb @ torch.diag_embed(wb**-0.5) @ vb.transpose(-1, -2)
            c = b_so @ a @ b_so
            w, v_ = _eigh(c, broadening_method, factor)
            v = b_so @ v_
        else:
            raise ValueError("Unknown scheme selected.")
    if sort_out and aux and b is not None:
        w, v = _eig_sort_out(w, v, True)
    elif sort_out and a.ndim > 2 and b is None:
        w, v = _eig_sort_out(w, v, True)
    return w, v
