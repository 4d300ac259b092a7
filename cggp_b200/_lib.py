"""ctypes binding of libcggp_b200.so (the C ABI in include/cggp_b200.h) + tensor plumbing.

PyTorch is used for device memory, streams and torch.distributed only; every compute call on this path goes
through the C ABI.  There is NO CPU fallback: importing the package without the built library, or creating a
context without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcggp_b200.so")

F32, F64 = 0, 1
SE, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
DIST_EUCLIDEAN, DIST_COVARIANCE, DIST_CORRELATION, DIST_SQEUCLIDEAN = 0, 1, 2, 3
OUT_KERNEL, OUT_DISTANCE = 0, 1
OP_DENSE, OP_SGPR = 0, 1
PRECOND_EYE, PRECOND_BLOCK, PRECOND_DENSE = 0, 1, 2
DISTANCE_CODES = {"euclidean": DIST_EUCLIDEAN, "covariance": DIST_COVARIANCE, "correlation": DIST_CORRELATION,
                  "sqeuclidean": DIST_SQEUCLIDEAN}


class CggpError(RuntimeError):
    pass


class Operator(C.Structure):
    """struct cggp_operator (include/cggp_b200.h)."""

    _fields_ = [
        ("struct_size", C.c_uint32), ("type", C.c_int32), ("dtype", C.c_int32), ("kind", C.c_int32),
        ("n", C.c_int64), ("dev_A", C.c_void_p), ("lda", C.c_int64),
        ("D", C.c_int32), ("variant", C.c_int32), ("variance", C.c_double), ("scale", C.c_double),
        ("dev_PX", C.c_void_p), ("dev_normsX", C.c_void_p), ("n_local", C.c_int64),
        ("dev_PZ", C.c_void_p), ("dev_normsZ", C.c_void_p), ("ldp", C.c_int64),
        ("dev_X32_big", C.c_void_p), ("dev_X32_small", C.c_void_p), ("dev_x32_norms", C.c_void_p),
        ("dev_Z32_big", C.c_void_p), ("dev_Z32_small", C.c_void_p), ("dev_z32_norms", C.c_void_p),
        ("tf32_nsplit", C.c_int32), ("_pad", C.c_int32),
    ]

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.struct_size = C.sizeof(Operator)


class Precond(C.Structure):
    """struct cggp_precond (include/cggp_b200.h)."""

    _fields_ = [
        ("type", C.c_int32), ("num_blocks", C.c_int32), ("block_size", C.c_int32), ("_pad", C.c_int32),
        ("dev_block_indices", C.c_void_p), ("dev_chol", C.c_void_p),
        ("dev_pinv", C.c_void_p), ("ldpinv", C.c_int64),
    ]


# name -> (restype, argtypes); tests/test_abi.py checks that every declaration of the header is exported
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
SIGNATURES = {
    "cggp_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "cggp_ctx_destroy": (_i, [_vp]),
    "cggp_ctx_set_stream": (_i, [_vp, _vp]),
    "cggp_last_error": (C.c_char_p, [_vp]),
    "cggp_launch_count": (_i64, [_vp]),
    "cggp_version": (C.c_char_p, []),
    "cggp_profile_enable": (_i, [_vp, _i]),
    "cggp_profile_read": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(_i64)]),
    "cggp_comm_unique_id": (_i, [_vp]),
    "cggp_ctx_comm_init": (_i, [_vp, _vp, _i, _i]),
    "cggp_ctx_comm_destroy": (_i, [_vp]),
    "cggp_peer_alloc": (_i, [_vp, _i64, _vp]),
    "cggp_peer_open": (_i, [_vp, _vp, _i]),
    "cggp_peer_close": (_i, [_vp]),
    "cggp_peer_enabled": (_i, [_vp]),
    "cggp_allreduce_sum": (_i, [_vp, _i, _vp, _i64]),
    "cggp_prepared_ld": (_i64, [_i]),
    "cggp_prepare_points": (_i, [_vp, _i, _vp, _i64, _i, _i64, C.POINTER(_d), _i, _vp, _i64, _vp]),
    "cggp_kernel_matrix": (_i, [_vp, _i, _i, _d, _i, _i, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i64, _d, _vp, _i64]),
    "cggp_kernel_matrix_backward": (_i, [_vp, _i, _i, _d, _vp, _i64, _vp, _i64, _i, _i64, C.POINTER(_d), _i, _vp, _i64,
                                          _vp, _vp]),
    "cggp_nearest_center": (_i, [_vp, _i, _i, _d, _i, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i64, _vp, _vp]),
    "cggp_cluster_stats": (_i, [_vp, _i, _vp, _vp, _i64, _i64, _vp, _vp]),
    "cggp_kuf_kfu_matvec": (_i, [_vp, _i, _i, _d, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i64, _vp, _i64, _i, _vp,
                                  _i64, _i]),
    "cggp_tf32_kp": (_i, [_i]),
    "cggp_tf32_rows": (_i64, [_i64]),
    "cggp_tf32_supported": (_i, [_vp, _i, _i]),
    "cggp_tf32_ring_plan": (_i, [_i, _i, _i, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int64)]),
    "cggp_tf32_sizes": (_i, [_i, _i64, _i, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cggp_tf32_prepare": (_i, [_vp, _i, _vp, _vp, _i64, _i, _i64, _vp, _vp, _vp]),
    "cggp_kuf_kfu_matvec_tf32": (_i, [_vp, _i, _d, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp,
                                       _i64, _i]),
    "cggp_kuf_times_tf32": (_i, [_vp, _i, _d, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64,
                                  _i]),
    "cggp_kuf_times": (_i, [_vp, _i, _i, _d, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i64, _vp, _i64, _i, _vp, _i64]),
    "cggp_kuf_gram": (_i, [_vp, _i, _i, _d, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i64, _vp, _i64, _i]),
    "cggp_symm_matmul": (_i, [_vp, _i, _vp, _i64, _i64, _vp, _i64, _i, _vp, _i64]),
    "cggp_block_cholesky": (_i, [_vp, _i, _vp, _i64, _i64, _vp, _i, _i, _vp]),
    "cggp_block_precond_apply": (_i, [_vp, _i, _i, _i64, _vp, C.POINTER(Precond), _vp]),
    "cggp_cg_fused_step": (_i, [_vp, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(Precond)]),
    "cggp_cg_solve": (_i, [_vp, C.POINTER(Operator), _vp, _vp, _i, _d, _i, _i, C.POINTER(Precond), _i, _vp,
                            C.POINTER(C.c_int32), _vp, _vp, _i64]),
    "cggp_predict_f": (_i, [_vp, _i, _i, _d, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i64, _vp, _i64, _vp, _d, _i, _i,
                            _vp, _vp, _vp, _vp, C.POINTER(C.c_int32)]),
    "cggp_elbo_terms": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _d, _vp]),
    "cggp_microbench": (_i, [_vp, _i, _i, C.POINTER(_d)]),
    "cggp_covertree_build": (_i, [_vp, _i, _vp, _i64, _i, _i64, _d, _i, _i, _i, C.POINTER(_vp)]),
    "cggp_covertree_destroy": (_i, [_vp]),
    "cggp_covertree_num_levels": (_i, [_vp]),
    "cggp_covertree_level_size": (_i64, [_vp, _i]),
    "cggp_covertree_level_radius": (_i, [_vp, _i, C.POINTER(_d)]),
    "cggp_covertree_level_points": (_i, [_vp, _vp, _i, _vp, _i64, C.POINTER(C.c_int32)]),
    "cggp_covertree_leaf_members": (_i, [_vp, _vp, _vp, _vp]),
    "cggp_covertree_cluster_stats": (_i, [_vp, _vp, _i, _vp, _i64, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def load_library():
    """Load libcggp_b200.so (no compute is done here; safe without a GPU)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CggpError(
                f"{LIB_PATH} is missing: build it with `python -m cggp_b200.build` "
                "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float64:
        return F64
    if dtype == torch.float32:
        return F32
    raise TypeError(f"cggp_b200 supports float32 / float64 tensors, got {dtype}")


class Context:
    """One cggp_ctx per (device): owns scratch memory, the device-side CG loop state and the NCCL communicator."""

    def __init__(self, device: int):
        if not torch.cuda.is_available():
            raise CggpError("cggp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device = int(device)
        handle = C.c_void_p()
        rc = self.lib.cggp_ctx_create(self.device, C.byref(handle))
        if rc != 0:
            raise CggpError("cggp_ctx_create failed: " + self.lib.cggp_last_error(None).decode())
        self.handle = handle
        self.world = 1
        self.rank = 0

    def check(self, rc: int):
        if rc != 0:
            raise CggpError(f"libcggp_b200 error {rc}: " + self.lib.cggp_last_error(self.handle).decode())

    def use_current_stream(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.check(self.lib.cggp_ctx_set_stream(self.handle, C.c_void_p(stream)))

    @property
    def launches(self) -> int:
        return int(self.lib.cggp_launch_count(self.handle))

    PROFILE_SECTIONS = ("kuf_kfu_matvec", "dense_symm_matmul", "cg_fused_step", "allreduce")

    def profile(self, on: bool = True):
        """Start (and clear) / stop per-section CUDA-event timing of the launches made through this context."""
        self.check(self.lib.cggp_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        """{section: (total_ms, launch_groups)}; synchronises the context stream."""
        out = {}
        for i, name in enumerate(self.PROFILE_SECTIONS):
            ms, cnt = C.c_double(0.0), C.c_int64(0)
            self.check(self.lib.cggp_profile_read(self.handle, i, C.byref(ms), C.byref(cnt)))
            out[name] = (ms.value, cnt.value)
        return out

    def init_comm(self, group=None):
        """Create the NCCL communicator of this rank from the torch.distributed process group (host plumbing):
        rank 0 makes the ncclUniqueId, the group broadcasts the 128 bytes, every rank calls cggp_ctx_comm_init."""
        import torch.distributed as dist

        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        buf = (C.c_char * 128)()
        if rank == 0:
            rc = self.lib.cggp_comm_unique_id(C.cast(buf, C.c_void_p))
            if rc != 0:
                raise CggpError("cggp_comm_unique_id failed: " + self.lib.cggp_last_error(None).decode())
        ident = [bytes(buf.raw)]
        dist.broadcast_object_list(ident, src=0, group=group)
        idbuf = (C.c_char * 128).from_buffer_copy(ident[0])
        self.check(self.lib.cggp_ctx_comm_init(self.handle, C.cast(idbuf, C.c_void_p), rank, world))
        self.world, self.rank = world, rank
        # NVLink peer buffers (ranks of one node): the CG loop's fused tail kernel all-reduces the partial product
        # straight out of peer memory (csrc/cg.cu cg_tail_kernel).  CGGP_PEER_ALLREDUCE=0 keeps everything on NCCL;
        # if the IPC mapping fails on any rank, all ranks agree to stay on NCCL.
        if world > 1 and os.environ.get("CGGP_PEER_ALLREDUCE", "1") == "1":
            self._init_peers(group, world)

    def _init_peers(self, group, world, slot_bytes: int = 1 << 20):
        """Peer buffers of the one-shot all-reduce (ranks of one node): exchange the CUDA IPC handles through the
        process group.  Every rank must end up in the same mode, so the outcome is agreed on before it is used."""
        import torch.distributed as dist

        handle = (C.c_char * 64)()
        ok = self.lib.cggp_peer_alloc(self.handle, slot_bytes, C.cast(handle, C.c_void_p)) == 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw) if ok else None, group=group)
        if ok and all(h is not None for h in handles):
            table = (C.c_char * (64 * world)).from_buffer_copy(b"".join(handles))
            ok = self.lib.cggp_peer_open(self.handle, C.cast(table, C.c_void_p), world) == 0
        else:
            ok = False
        agreed = [None] * world
        dist.all_gather_object(agreed, bool(ok), group=group)
        if not all(agreed):
            self.lib.cggp_peer_close(self.handle)

    @property
    def peer_allreduce(self) -> bool:
        return bool(self.lib.cggp_peer_enabled(self.handle))

    def allreduce_sum_(self, t: torch.Tensor):
        self.use_current_stream()
        self.check(self.lib.cggp_allreduce_sum(self.handle, dtype_code(t.dtype), ptr(t), t.numel()))
        return t

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.cggp_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_contexts = {}


def context(device=None) -> Context:
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def as_device_tensor(x, dtype=None, device=None) -> torch.Tensor:
    """Zero-copy import of a device tensor: torch tensors pass through, anything exposing ``__dlpack__``
    (TensorFlow eager tensors via ``tf.experimental.dlpack``, CuPy, JAX) is wrapped with ``torch.from_dlpack``;
    host arrays are uploaded once (H2D copy, plumbing)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, "__dlpack__") and not type(x).__module__.startswith("numpy"):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(x)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise CggpError("cggp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        t = t.to(torch.device("cuda", torch.cuda.current_device()) if device is None else device)
    return t


def row_major(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor with unit inner stride (leading dimension = stride(0)); copies only if needed."""
    if t.dim() != 2:
        raise ValueError(f"expected a matrix, got shape {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        return t.contiguous()
    if t.shape[1] == 1 and t.stride(0) < 1:
        return t.contiguous()
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        return t.contiguous()
    return t
