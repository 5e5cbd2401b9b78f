"""ctypes binding of libqec_b200.so (the C ABI declared in include/qec_b200.h).

There is no CPU fallback and no alternative backend: if the library is missing
the import of any op fails loudly."""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libqec_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_LL = ctypes.c_longlong
_F = ctypes.c_float
_D = ctypes.c_double

# name -> (argtypes, kernels launched per call)
SIGNATURES = {
    "qec_b200_version": ([], 0),
    # (mean LRF + point rotation: one launch for K <= 32 and for the tile kernel (I >= 32); two otherwise)
    "qec_lrf_mean_fwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _P], lambda a: 1 if (a[6] <= 32 or a[7] >= 32 or not a[4]) else 2),
    "qec_lrf_mean_bwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P], 1),
    "qec_votes_fwd": ([_P, _P, _P, _I, _I, _I, _I, _P], 1),
    "qec_votes_bwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _P], 1),
    "qec_routing_fwd": ([_P] * 8 + [_I, _I, _I, _I, _P], 1),
    "qec_routing_bwd": ([_P] * 11 + [_I, _I, _I, _I, _P], 1),
    "qec_routing_fused_supported": ([_I, _I, _I, _I], 0),
    "qec_routing_fused_fwd": ([_P] * 9 + [_I, _I, _I, _I, _I, _P], 1),
    "qec_routing_fused_bwd": ([_P] * 13 + [_I, _I, _I, _I, _I, _P], 1),
    "qec_routing_fused_packed_supported": ([_I, _I, _I, _I], 0),
    "qec_routing_fused_set_wide": ([_I], 0),
    "qec_routing_fused_set_pp": ([_I], 0),
    "qec_routing_fused_pp_supported": ([_I, _I, _I, _I], 0),
    "qec_routing_fused_pp_fwd": ([_P] * 9 + [_I, _I, _I, _I, _I, _P], 1),
    "qec_routing_fused_bwd_split": ([_P] * 14 + [_I, _I, _I, _I, _I, _P], 1),
    "qec_routing_fused_fwd_scalar": ([_P] * 9 + [_I, _I, _I, _I, _I, _P], 1),
    "qec_routing_fused_bwd_scalar": ([_P] * 13 + [_I, _I, _I, _I, _I, _P], 1),
    "qec_routing_small_supported": ([_I, _I, _I, _I], 0),
    "qec_routing_small_fwd": ([_P] * 11 + [_I, _I, _I, _I, _P], 1),
    "qec_routing_small_bwd": ([_P] * 14 + [_I, _I, _I, _I, _P], 1),
    "qec_symeig4_fwd": ([_P, _P, _P, _P, _LL, _P], 1),
    "qec_symeig4_bwd": ([_P, _P, _P, _P, _P, _LL, _I, _P], 1),
    "qec_gather": ([_P] * 9 + [_I, _I, _I, _I, _I, _P], 1),
    "qec_knn": ([_P] * 5 + [_I, _I, _I, _I, _P], 1),
    "qec_split_operand": ([_P, _P, _F, _P, _P, _P, _LL, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], 1),
    "qec_fold_chunks": ([_P, _P, _P, _P, _P, _I, _LL, _I, _I, _I, _P], 1),
    "qec_gen2_bwd_supported": ([_I, _I, _I], 0),
    "qec_gen2_bwd_workspace": ([_I, _I, _I], 0),
    "qec_gen2_bwd": ([_P] * 8 + [_I, _I, _I, _I, _P], lambda a: 2 + (1 if a[4] else 0) + (1 if a[5] else 0)),
    "qec_spread_loss": ([_P, _P, _F, _P, _P, _P, _P, _I, _I, _P], 2),
    "qec_gen2_fwd_supported": ([_I, _I, _I], 0),
    "qec_gen2_fwd_workspace": ([_I, _I, _I], 0),
    "qec_gen2_fwd": ([_P] * 6 + [_I, _I, _I, _P], 2),
    "qec_linear_supported": ([_I, _I], 0),
    "qec_linear_bwd_workspace": ([_I, _I, _I], 0),
    "qec_linear_fwd": ([_P] * 4 + [_I, _I, _I, _P], 1),
    "qec_linear_bwd": ([_P] * 7 + [_I, _I, _I, _P], lambda a: 1 + (1 if (a[4] or a[5]) else 0)),
    "qec_compose_supported": ([_I], 0),
    "qec_compose_fwd": ([_P] * 6 + [_I, _I, _I, _P], 1),
    "qec_compose_bwd": ([_P] * 9 + [_I, _I, _I, _P], 1),
    "qec_class_pose_loss": ([_P, _P, _F, _P, _F, _P, _P, _P, _P, _I, _I, _I, _P], 1),
    "qec_quat_eval": ([_P, _P, _P, _P, _LL, _P], 1),
    "qec_adam_flat": ([_P, _P, _P, _P, _LL, _F, _D, _D, _F, _F, _F, _I, _P, _P], 1),
    "qec_ffma_peak": ([_P, _I, _I, _P], 1),
    "qec_pipe_probe": ([_P, _I, _I, _I, _P], 1),
}

_lib = None
_lock = threading.Lock()
_launches = 0
_timing = None  # list of (name, int-args, start event, end event) while kernel timing is on


class QecLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library once (thread-safe: DataParallel calls from one thread per GPU)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise QecLibraryError(
                    "%s not found: build it with `python -m qecnetworks_b200.build` "
                    "(there is no CPU or PyTorch fallback for the QEC routing kernels)" % LIB_PATH
                )
            lib = ctypes.CDLL(LIB_PATH)
            for name, (argtypes, _) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = argtypes
                fn.restype = ctypes.c_longlong if name.endswith("_workspace") else ctypes.c_int
            _lib = lib
    return _lib


def call(name, *args):
    """Invoke an entry point; raise on a non-zero status."""
    global _launches
    fn = getattr(load(), name)
    if _timing is not None:
        import torch

        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()  # on torch's current stream == the stream the kernels are launched on
        rc = fn(*args)
        e1.record()
        _timing.append((name, tuple(a for a in args[:-1] if isinstance(a, int) and a < (1 << 31)), e0, e1))
    else:
        rc = fn(*args)
    if rc != 0:
        if rc < 0:
            raise QecLibraryError("%s: argument %d is invalid" % (name, -rc))
        raise QecLibraryError("%s: CUDA error %d at launch" % (name, rc))
    n = SIGNATURES[name][1]
    _launches += n(args) if callable(n) else n


def launch(name, *args):
    """call() for tensor arguments: every torch tensor among `args` is passed as its device pointer (None as a
    null pointer), all of them must live on ONE CUDA device, and the kernels are launched on THAT device's
    current stream with that device made current for the call -- like a native PyTorch op, whatever
    torch.cuda.current_device() happens to be (a module moved with net.to('cuda:1'), or the reference's
    nn.DataParallel with one thread per GPU).  The stream is appended as the last argument."""
    import torch

    dev = None
    raw = []
    for i, a in enumerate(args):
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                raise QecLibraryError("%s: argument %d is a CPU tensor (the QEC kernels have no CPU implementation)"
                                      % (name, i + 1))
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise QecLibraryError("%s: argument %d lives on %s, the others on %s" % (name, i + 1, a.device, dev))
            raw.append(a.data_ptr())
        else:
            raw.append(a)
    if dev is None:
        raise QecLibraryError("%s: no tensor argument to take the device from" % name)
    with torch.cuda.device(dev):
        call(name, *raw, torch.cuda.current_stream(dev).cuda_stream)


def query(name, *args):
    """entry points that return a value rather than a status"""
    return getattr(load(), name)(*args)


def launches():
    """Number of kernels of this library launched by this process so far."""
    return _launches


def start_kernel_timing():
    """record CUDA events around every entry-point call (bench.py's per-kernel roofline pass)"""
    global _timing
    _timing = []


def stop_kernel_timing():
    """-> {(name, dims): [ms, ...]}; the caller must have synchronised the device"""
    global _timing
    rec, _timing = _timing, None
    out = {}
    for name, dims, e0, e1 in rec or []:
        out.setdefault((name, dims), []).append(e0.elapsed_time(e1))
    return out


def ptr(t):
    return None if t is None else t.data_ptr()
