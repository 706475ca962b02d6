"""Device context and value packing for the host-side mirror of the reference API.

A field element travels as 16 little-endian bytes (Rust's in-memory u128); a vector of
them is a numpy uint64 array of shape (n, 2) = (lo, hi) on the host, or a torch int64
CUDA tensor of the same shape when it should stay in HBM between calls.  Every function
of the mirror accepts a list of Python ints, a numpy array or a CUDA tensor, and returns
the same kind it was given (lists of ints for lists).

There is no CPU path: creating a Context without a CUDA device raises.
"""
import ctypes

import numpy as np

from . import _lib

P = 1 + 407 * (1 << 119)                                   # src/field/field.rs:9-10
_M64 = (1 << 64) - 1

ERRORS = {-1: "CUDA", -2: "ARG", -3: "EMPTY", -4: "NOT_POW2", -5: "TOO_LONG", -6: "ROOT_ORDER", -7: "DIV_ZERO",
          -8: "INDEX", -9: "LENGTH", -10: "ROUNDS", -11: "DEGREE", -12: "CALLBACK"}


class ZkbError(RuntimeError):
    """Raised where the reference would panic; .code is the ZKB_ERR_* value."""

    def __init__(self, code, message):
        super().__init__("%s (ZKB_ERR_%s)" % (message, ERRORS.get(code, code)))
        self.code = code


def le16(v):
    return (ctypes.c_uint8 * 16).from_buffer_copy(int(v).to_bytes(16, "little"))


def from_le16(buf):
    return int.from_bytes(bytes(buf), "little")


def pack(vals):
    """list of ints -> (n, 2) uint64"""
    a = np.empty((len(vals), 2), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i, 0] = v & _M64
        a[i, 1] = v >> 64
    return a


def unpack(arr):
    arr = np.ascontiguousarray(arr).view(np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in arr.tolist()]


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class Vec:
    """A borrowed view of caller data as (pointer, n, kind)."""

    def __init__(self, x):
        self.keep = x
        if _is_tensor(x):
            if not x.is_cuda:
                x = x.contiguous().numpy()
            else:
                assert x.is_contiguous() and x.dim() == 2 and x.shape[1] == 2 and x.element_size() == 8, \
                    "device vectors are contiguous (n, 2) 64-bit tensors"
                self.kind, self.n, self.ptr = "cuda", x.shape[0], x.data_ptr()
                self.device = x.device
                return
        if isinstance(x, np.ndarray):
            a = np.ascontiguousarray(x).view(np.uint64).reshape(-1, 2)
            self.kind = "numpy"
        else:
            a = pack(list(x))
            self.kind = "list"
        self.keep = a
        self.n, self.ptr = a.shape[0], a.ctypes.data


class Context:
    """One per GPU (zkb_ctx).  `stream`: a raw cudaStream_t; "torch" (default) = torch's current
    stream on that device, so calls on CUDA tensors are ordered with the caller's torch work;
    "own" = a private non-blocking stream (the caller must ctx.sync() before touching results
    from another stream)."""

    def __init__(self, device=0, stream="torch"):
        self.lib = _lib.lib()
        self.device = device
        if stream == "torch":
            stream = None
            try:
                import torch
                if torch.cuda.is_available():
                    # torch's default stream is the legacy NULL stream: its explicit handle is
                    # cudaStreamLegacy (0x1); a NULL argument would mean "create a private stream"
                    stream = torch.cuda.current_stream(device).cuda_stream or 1
            except ImportError:
                pass
        elif stream == "own":
            stream = None
        h = ctypes.c_void_p()
        rc = self.lib.zkb_ctx_create(device, ctypes.c_void_p(stream) if stream else None, ctypes.byref(h))
        if rc != 0 or not h.value:
            raise ZkbError(rc, "zkb_ctx_create(device=%d) failed: no usable CUDA device; there is no CPU fallback" % device)
        self.h = h

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.zkb_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise ZkbError(rc, self.lib.zkb_last_error(self.h).decode())

    def sync(self):
        self.check(self.lib.zkb_ctx_sync(self.h))

    @property
    def launches(self):
        return int(self.lib.zkb_ctx_launches(self.h))

    def profile(self, enable=True, reset=False):
        """Per-kernel-class CUDA-event timing of every launch on this context's stream."""
        self.check(self.lib.zkb_ctx_profile(self.h, (2 if reset else 1) if enable else 0))

    def profile_read(self):
        """{kernel name: (total ms, launches)} accumulated since the last reset."""
        out, i = {}, 0
        while True:
            name = self.lib.zkb_kernel_name(i)
            if not name:
                return out
            ms, cnt = ctypes.c_double(0), ctypes.c_uint64(0)
            self.check(self.lib.zkb_ctx_profile_read(self.h, i, ctypes.byref(ms), ctypes.byref(cnt)))
            if cnt.value:
                out[name.decode()] = (ms.value, int(cnt.value))
            i += 1

    def out_like(self, v, n):
        """An output buffer of n elements of the same kind as the input view `v`."""
        if v.kind == "cuda":
            import torch
            t = torch.empty((n, 2), dtype=torch.int64, device=v.device)
            return t, t.data_ptr()
        a = np.empty((n, 2), dtype=np.uint64)
        return a, a.ctypes.data

    @staticmethod
    def finish(v, out):
        return unpack(out) if v.kind == "list" else out


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
