"""Drop-in for the reference's ``torch_autograd_solver.symeig`` on the hot-path shape
([n,4,4] float32 CUDA): eigenvalues ascending, eigenvectors as ROWS of V
(models/pytorch-autograd-solver/torch_autograd_solver/__init__.py:141-159, 68-98).

With this on ``sys.modules['torch_autograd_solver']`` the UNMODIFIED reference
QecModule runs on a B200 with our eigensolver in place of pytorch-cusolver."""
import torch

from ._lib import launch, QecLibraryError


class BatchSymeig4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, fold):
        n = a.shape[0]
        a = a.contiguous()
        w = torch.empty(n, 4, device=a.device, dtype=torch.float32)
        V = torch.empty(n, 4, 4, device=a.device, dtype=torch.float32)
        status = torch.zeros(1, device=a.device, dtype=torch.int32)
        if n:
            launch("qec_symeig4_fwd", a, w, V, status, n)
        ctx.save_for_backward(w, V)
        ctx.fold = int(fold)
        ctx.status = status  # device-side count of non-converged matrices; never synced here
        return w, V

    @staticmethod
    def backward(ctx, gw, gV):
        w, V = ctx.saved_tensors
        n = w.shape[0]
        gM = torch.zeros(n, 4, 4, device=w.device, dtype=torch.float32)
        if n and (gw is not None or gV is not None):
            gw = None if gw is None else gw.contiguous()
            gV = None if gV is None else gV.contiguous()
            launch("qec_symeig4_bwd", gw, gV, w, V, gM, n, ctx.fold)
        return gM, None


def symeig(a, b=None, fold=True, **kwargs):
    """symeig(a) -> (w, V).  Only the batched 4x4 float32 CUDA case of the reference's
    dispatcher is implemented (the only one the QEC-Net hot path uses); everything else
    raises instead of silently falling back."""
    if b is not None:
        raise NotImplementedError("generalized symeig is dead code in the reference and not provided")
    if a.dim() != 3 or a.shape[-1] != 4 or a.shape[-2] != 4:
        raise NotImplementedError("only [n,4,4] input is supported, got %s" % (tuple(a.shape),))
    if not a.is_cuda or a.dtype != torch.float32:
        raise QecLibraryError("symeig needs a float32 CUDA tensor (no CPU fallback)")
    return BatchSymeig4.apply(a, fold)
