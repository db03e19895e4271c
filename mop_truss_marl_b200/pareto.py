"""Batched Pareto-front bookkeeping on the GPU (C ABI in ``include/tpareto.h``): ``utils.simple_cull`` and
``utils.union_rectangles_fastest`` of the reference (``test/*/code/utils.py:11-217, 463-530``) for B environments at
once.  PyTorch only provides the device buffers and the stream."""
from __future__ import annotations

import ctypes as C

import torch

from . import capi

PARETO_EXPORTS = ("tpareto_last_error", "tpareto_front_hv", "tpareto_front_hv_thin", "tpareto_state_data")
MAX_POINTS = 256

_lib = capi.lib
_lib.tpareto_last_error.restype = C.c_char_p
_lib.tpareto_front_hv.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 8
_lib.tpareto_front_hv_thin.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 5
_lib.tpareto_state_data.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def front_hv(points: torch.Tensor, counts: torch.Tensor | None = None, ref_point=(1.0, 1.0), thin_pick: torch.Tensor | None = None,
             max_front: int = 50):
    """points [B,P,4] float32 CUDA (obj1, obj2, con1, con2), counts [B] int32 (valid points per environment).
    Returns ``front_idx`` [B,P] (front members in the reference's order, -1 padded), ``front_len`` [B],
    ``stats`` [B,5] = (max_distance, dis_distance, p_norm_inv_cd, sum_distance, std_cd) and ``hv`` [B].
    ``thin_pick`` [B, max_front - 2] int32 CUDA: the draw ``random.sample(range(F - 2), max_front - 2)`` per environment,
    applied to fronts of more than ``max_front`` members like ``utils.simple_cull`` does (:104-131); without it larger
    fronts stay unthinned."""
    if not (points.is_cuda and points.dtype == torch.float32 and points.dim() == 3 and points.shape[2] == 4
            and points.is_contiguous()):
        raise ValueError("points must be a contiguous float32 CUDA tensor [B,P,4]")
    B, P = int(points.shape[0]), int(points.shape[1])
    if counts is not None and not (counts.is_cuda and counts.dtype == torch.int32 and tuple(counts.shape) == (B,)):
        raise ValueError("counts must be an int32 CUDA tensor [B]")
    dev = points.device
    front_idx = torch.empty(B, P, dtype=torch.int32, device=dev)
    front_len = torch.empty(B, dtype=torch.int32, device=dev)
    stats = torch.empty(B, 5, dtype=torch.float64, device=dev)
    hv = torch.empty(B, dtype=torch.float64, device=dev)
    ref = (C.c_double * 2)(float(ref_point[0]), float(ref_point[1]))
    if thin_pick is not None and not (thin_pick.is_cuda and thin_pick.dtype == torch.int32 and thin_pick.is_contiguous()
                                      and tuple(thin_pick.shape) == (B, max_front - 2)):
        raise ValueError("thin_pick must be a contiguous int32 CUDA tensor [B, max_front - 2]")
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if thin_pick is None:
            rc = _lib.tpareto_front_hv(B, P, _ptr(points), _ptr(counts), C.cast(ref, C.c_void_p), _ptr(front_idx),
                                       _ptr(front_len), _ptr(stats), _ptr(hv), st)
        else:
            rc = _lib.tpareto_front_hv_thin(B, P, _ptr(points), _ptr(counts), C.cast(ref, C.c_void_p), _ptr(thin_pick),
                                            int(max_front), _ptr(front_idx), _ptr(front_len), _ptr(stats), _ptr(hv), st)
    if rc != 0:
        raise capi.TfemError("libtfem pareto error %d: %s" % (rc, _lib.tpareto_last_error().decode()))
    return {"front_idx": front_idx, "front_len": front_len, "stats": stats, "hv": hv}


def state_data(points: torch.Tensor, front_idx: torch.Tensor | None = None, front_len: torch.Tensor | None = None,
               index: torch.Tensor | None = None, rows: int = 50, max_front: int = 50):
    """``pareto_state_data`` (``truss2D_ENV.py:22-41``) for B fronts, padded / cut to ``rows`` rows like the driver does
    (``master_DDPG_truss2D_MO.py:499-517``).  ``points`` [B,P,4] float32 CUDA; ``front_idx`` [B,P] / ``front_len`` [B] as
    returned by :func:`front_hv` (None: the points are the front, in order); ``index`` [B] int32: the front member whose
    state is being built; ``max_front``: the ENV module's ``MAX_FRONT`` (50 in test/, 20 in train/).  Returns ``x_p`` [B,rows,4], ``A_p`` [B,rows,rows]: the actor's Pareto-graph inputs."""
    if not (points.is_cuda and points.dtype == torch.float32 and points.dim() == 3 and points.shape[2] == 4
            and points.is_contiguous()):
        raise ValueError("points must be a contiguous float32 CUDA tensor [B,P,4]")
    B, P = int(points.shape[0]), int(points.shape[1])
    dev = points.device
    for name, t, shape in (("front_idx", front_idx, (B, P)), ("front_len", front_len, (B,)), ("index", index, (B,))):
        if t is not None and not (t.is_cuda and t.dtype == torch.int32 and tuple(t.shape) == shape and t.is_contiguous()):
            raise ValueError("%s must be a contiguous int32 CUDA tensor of shape %s" % (name, shape))
    x_p = torch.empty(B, rows, 4, dtype=torch.float32, device=dev)
    A_p = torch.empty(B, rows, rows, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.tpareto_state_data(B, P, int(rows), int(max_front), _ptr(points), _ptr(front_idx), _ptr(front_len), _ptr(index), _ptr(x_p),
                                     _ptr(A_p), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise capi.TfemError("libtfem pareto error %d: %s" % (rc, _lib.tpareto_last_error().decode()))
    return x_p, A_p
