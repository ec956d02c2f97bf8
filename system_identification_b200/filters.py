"""The filters of the reference's read_data (reference demo/solo_identification.py:15-32) on the device.

    b, a = butter_lowpass(5, 0.15)                    # scipy.signal.butter(5, 0.15, btype='low', analog=False)
    dq_f = filtfilt(b, a, dq)                         # scipy.signal.filtfilt(b, a, dq, axis=1)
    dq_s = savgol_filter(dq, 21, 5)                   # scipy.signal.savgol_filter(dq, 21, 5)

Arrays are channel-major (channels, N) CUDA float64 tensors, exactly what the fused kernel consumes next; the work is
done by sysid_filtfilt / sysid_savgol in libsysid_b200.so (hand-written kernels, no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _require_cuda, _stream


def butter_lowpass(order, wn):
    """Digital Butterworth low-pass (b, a): analog prototype -> frequency pre-warp -> bilinear transform, the textbook
    design scipy.signal.butter(order, wn, btype='low', analog=False) implements (wn relative to the Nyquist frequency)."""
    order = int(order)
    if order < 1 or not (0.0 < wn < 1.0):
        raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
    k = np.arange(-order + 1, order, 2)
    p = -np.exp(1j * np.pi * k / (2 * order))                 # unit-circle poles of the analog prototype
    warped = 4.0 * np.tan(np.pi * wn / 2.0)                   # pre-warp, sampling rate 2
    p = p * warped
    gain = warped ** order
    fs2 = 4.0
    pd = (fs2 + p) / (fs2 - p)                                # bilinear transform; the order zeros at infinity map to z = -1
    gd = gain * np.real(1.0 / np.prod(fs2 - p))
    b = gd * np.poly(-np.ones(order))
    a = np.real(np.poly(pd))
    return b.astype(np.float64), a.astype(np.float64)


def _check(x, name):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.stride(1) == 1):
        raise ValueError(f"{name}: expected a CUDA float64 tensor (channels, N) with unit inner stride")
    return x


def filtfilt(b, a, x, out=None, float32_input=False):
    """scipy.signal.filtfilt(b, a, x, axis=1) with its defaults (odd extension, padlen 3*max(len(a), len(b))).
    float32_input=True: x holds widened float32 values (the reference filters np.loadtxt(dtype=float32) arrays); scipy
    then builds the odd extension in float32, which is reproduced so that the result matches the reference's exactly."""
    _require_cuda()
    lib = _lib.load()
    x = _check(x, "x")
    b = np.ascontiguousarray(np.atleast_1d(b), dtype=np.float64)
    a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)
    ch, N = x.shape
    padlen = 3 * max(len(a), len(b))
    if N <= padlen:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {padlen}.")
    if out is None:
        out = torch.empty((ch, N), dtype=torch.float64, device=x.device)
    out = _check(out, "out")
    if out.shape != x.shape or out.stride(0) != x.stride(0):
        if out.shape != x.shape:
            raise ValueError("out: shape must match x")
        x = x.contiguous() if x.stride(0) != N else x
        if out.stride(0) != x.stride(0):
            raise ValueError("out and x must share the leading dimension")
    nbytes = lib.sysid_filtfilt_workspace_bytes(ch, N, max(len(a), len(b)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _lib.check(lib.sysid_filtfilt(b.ctypes.data_as(C.c_void_p), len(b), a.ctypes.data_as(C.c_void_p), len(a), _ptr(x), _ptr(out),
                                  ch, N, x.stride(0), 1 if float32_input else 0, _ptr(ws), ws.numel(), _stream()))
    return out


def savgol_filter(x, window_length, polyorder, float32_input=False):
    """scipy.signal.savgol_filter(x, window_length, polyorder) with its defaults (deriv=0, mode='interp'), along axis 1.
    float32_input=True rounds the result through float32, which is the dtype scipy returns for a float32 log (the reference's
    np.loadtxt(dtype=float32) arrays, demo/solo_identification.py:26-32) and therefore what the reference hands on to
    pinocchio.  scipy's float32 path is not bit-reproducible across versions (ndimage accumulates in double, but 1.15+ runs
    the polynomial edge fit in float32): agreement with it is to float32 precision (3e-5 relative on the test fixture)."""
    _require_cuda()
    lib = _lib.load()
    x = _check(x, "x")
    ch, N = x.shape
    if polyorder >= window_length:
        raise ValueError("polyorder must be less than window_length.")
    if window_length > N:
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
    x = x if x.stride(0) == N else x.contiguous()
    out = torch.empty((ch, N), dtype=torch.float64, device=x.device)
    ws = torch.empty(lib.sysid_savgol_workspace_bytes(int(window_length)), dtype=torch.uint8, device=x.device)
    _lib.check(lib.sysid_savgol(int(window_length), int(polyorder), _ptr(x), _ptr(out), ch, N, x.stride(0), _ptr(ws), ws.numel(), _stream()))
    return out.to(torch.float32).to(torch.float64) if float32_input else out


def preprocess(robot_q, robot_dq, robot_ddq, robot_tau, robot_contact, filter_type):
    """The tail of the reference's read_data: float32 logs in (as np.loadtxt(dtype=float32) returns them), filtered fp64
    device arrays out.  filter_type: "butterworth" (order 5, cutoff 0.15), "savitzky" (window 21, order 5), anything
    else: no filtering (the reference then keeps float32; here the values are widened exactly)."""
    from .ops import to_device
    q, dq, ddq, tau, cnt = (to_device(np.asarray(a)) for a in (robot_q, robot_dq, robot_ddq, robot_tau, robot_contact))
    if filter_type == "butterworth":
        b, a = butter_lowpass(5, 0.15)
        f32 = all(np.asarray(v).dtype == np.float32 for v in (robot_dq, robot_ddq, robot_tau))
        dq, ddq, tau = (filtfilt(b, a, v, float32_input=f32) for v in (dq, ddq, tau))
    elif filter_type == "savitzky":
        f32 = all(np.asarray(v).dtype == np.float32 for v in (robot_dq, robot_ddq, robot_tau))
        dq, ddq, tau = (savgol_filter(v, 21, 5, float32_input=f32) for v in (dq, ddq, tau))
    return q, dq, ddq, tau, cnt
