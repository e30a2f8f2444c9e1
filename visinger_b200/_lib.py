"""ctypes binding of the C-ABI library (include/visinger_b200.h).

The library is built IN-TREE with nvcc for sm_100a (`visinger_b200/lib/libvisinger_b200.so`) and
loaded with ctypes: no torch types cross the boundary, only raw device pointers, sizes and the
CUDA stream handle.  PyTorch is used for device memory and streams only.  There is no CPU or
eager fallback: if the library cannot be built or loaded, or a tensor is not on a CUDA device,
the call raises.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import shutil
import subprocess
import threading
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIBDIR = os.path.join(_HERE, "lib")
_LIBPATH = os.path.join(_LIBDIR, "libvisinger_b200.so")
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "visinger_b200.h")
_SOURCES = ["api.cu", "pack.cu", "run_f32.cu", "run_tc.cu", "output.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-diag-suppress", "550"]

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_BF16X3 = 2
_PRECISIONS = {"fp32": PRECISION_FP32, "float32": PRECISION_FP32, "bf16": PRECISION_BF16, "bfloat16": PRECISION_BF16,
               "bf16x3": PRECISION_BF16X3}

VSG_MAX_UPS = 8
VSG_MAX_RESBLOCK_KERNELS = 4
VSG_MAX_RESBLOCK_DILATIONS = 4


def precision_code(p) -> int:
    if isinstance(p, int):
        return p
    try:
        return _PRECISIONS[str(p).lower()]
    except KeyError:
        raise ValueError(f"unknown precision {p!r}; use 'fp32', 'bf16' or 'bf16x3'") from None


class VsgConfig(ctypes.Structure):
    _fields_ = [
        ("flow_channels", ctypes.c_int32), ("flow_hidden", ctypes.c_int32), ("flow_kernel_size", ctypes.c_int32),
        ("flow_dilation_rate", ctypes.c_int32), ("flow_n_layers", ctypes.c_int32), ("flow_n_flows", ctypes.c_int32),
        ("flow_gin", ctypes.c_int32),
        ("dec_initial_channel", ctypes.c_int32), ("dec_resblock", ctypes.c_int32), ("dec_n_kernels", ctypes.c_int32),
        ("dec_resblock_kernel_sizes", ctypes.c_int32 * VSG_MAX_RESBLOCK_KERNELS),
        ("dec_n_dilations", ctypes.c_int32 * VSG_MAX_RESBLOCK_KERNELS),
        ("dec_resblock_dilations", (ctypes.c_int32 * VSG_MAX_RESBLOCK_DILATIONS) * VSG_MAX_RESBLOCK_KERNELS),
        ("dec_n_ups", ctypes.c_int32),
        ("dec_upsample_rates", ctypes.c_int32 * VSG_MAX_UPS),
        ("dec_upsample_kernel_sizes", ctypes.c_int32 * VSG_MAX_UPS),
        ("dec_upsample_initial_channel", ctypes.c_int32), ("dec_gin", ctypes.c_int32),
    ]


class VsgEncConfig(ctypes.Structure):
    _fields_ = [("in_channels", ctypes.c_int32), ("out_channels", ctypes.c_int32), ("hidden_channels", ctypes.c_int32),
                ("kernel_size", ctypes.c_int32), ("dilation_rate", ctypes.c_int32), ("n_layers", ctypes.c_int32),
                ("gin_channels", ctypes.c_int32)]


class VsgRelEncConfig(ctypes.Structure):
    _fields_ = [("hidden_channels", ctypes.c_int32), ("filter_channels", ctypes.c_int32), ("n_heads", ctypes.c_int32),
                ("n_layers", ctypes.c_int32), ("kernel_size", ctypes.c_int32), ("window_size", ctypes.c_int32),
                ("gin_channels", ctypes.c_int32)]


class VsgTensor(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.c_void_p), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int64 * 4)]


def _source_files():
    files = [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cu", ".cuh", ".h", ".inc"))]
    return files + [_HEADER]


def _source_hash() -> str:
    h = hashlib.sha256()
    for f in _source_files():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _nvcc() -> Optional[str]:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else shutil.which("nvcc")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into the in-tree shared library (no GPU needed)."""
    os.makedirs(_LIBDIR, exist_ok=True)
    want = _source_hash()
    stamp = _LIBPATH + ".hash"
    if not force and os.path.exists(_LIBPATH) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return _LIBPATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("visinger_b200: libvisinger_b200.so is missing or stale and nvcc was not found; "
                           "there is no fallback path")
    # one nvcc per translation unit, in parallel, then one link: run_tc.cu alone is most of the serial build time
    objdir = os.path.join(os.path.dirname(_HERE), "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(_CSRC, src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, proc

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(_SOURCES)) as pool:
        results = list(pool.map(compile_one, _SOURCES))
    for src, _, proc in results:
        if proc.returncode != 0:
            raise RuntimeError(f"visinger_b200: nvcc failed on {src}\n" + proc.stdout + proc.stderr)
        if verbose:
            print(proc.stdout + proc.stderr, flush=True)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", _LIBPATH + ".tmp"] + [o for _, o, _ in results]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("visinger_b200: link failed\n" + proc.stdout + proc.stderr)
    os.replace(_LIBPATH + ".tmp", _LIBPATH)
    with open(stamp, "w") as fh:
        fh.write(want)
    return _LIBPATH


_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Load (building first if stale) the C-ABI library.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        # VSG_LIB_OVERRIDE: load another build of the same library (A/B timing of two kernel versions on one GPU box)
        path = os.environ.get("VSG_LIB_OVERRIDE") or build()
        L = ctypes.CDLL(path)
        vp, i32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t
        L.vsg_abi_version.restype = ctypes.c_int
        L.vsg_last_error.restype = ctypes.c_char_p
        L.vsg_last_launch_count.restype = i32
        L.vsg_pack_create.restype = ctypes.c_int
        L.vsg_pack_create.argtypes = [ctypes.POINTER(VsgConfig), ctypes.POINTER(VsgTensor), i32, ctypes.c_char_p,
                                      ctypes.c_char_p, i32, ctypes.POINTER(vp)]
        L.vsg_pack_destroy.restype = None
        L.vsg_pack_destroy.argtypes = [vp]
        L.vsg_workspace_bytes.restype = sz
        L.vsg_workspace_bytes.argtypes = [vp, i32, i32, i32]
        L.vsg_hop_size.restype = i32
        L.vsg_hop_size.argtypes = [vp]
        L.vsg_prior_sample.restype = ctypes.c_int
        L.vsg_prior_sample.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
        L.vsg_flow_forward.restype = ctypes.c_int
        L.vsg_flow_forward.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]
        L.vsg_generator_forward.restype = ctypes.c_int
        L.vsg_generator_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]
        L.vsg_infer.restype = ctypes.c_int
        L.vsg_infer.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]
        L.vsg_wav_to_int16.restype = ctypes.c_int
        L.vsg_wav_to_int16.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
        L.vsg_debug_conv1d_bf16.restype = ctypes.c_int
        L.vsg_debug_conv1d_bf16.argtypes = [vp, vp, vp, vp, vp, ctypes.c_float, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32,
                                            i32]
        L.vsg_set_tc_options.restype = ctypes.c_int
        L.vsg_set_tc_options.argtypes = [i32, i32, i32, i32]
        L.vsg_debug_pair_bf16.restype = ctypes.c_int
        L.vsg_debug_pair_bf16.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_float, vp, vp, vp, i32, i32, i32, i32, i32, i32]
        L.vsg_enc_pack_create.restype = ctypes.c_int
        L.vsg_enc_pack_create.argtypes = [ctypes.POINTER(VsgEncConfig), ctypes.POINTER(VsgTensor), i32, ctypes.c_char_p, i32,
                                          ctypes.POINTER(ctypes.c_void_p)]
        L.vsg_posterior_workspace_bytes.restype = ctypes.c_size_t
        L.vsg_posterior_workspace_bytes.argtypes = [vp, i32, i32, i32]
        L.vsg_posterior_forward.restype = ctypes.c_int
        L.vsg_posterior_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, ctypes.c_size_t, vp]
        L.vsg_relenc_pack_create.restype = ctypes.c_int
        L.vsg_relenc_pack_create.argtypes = [ctypes.POINTER(VsgRelEncConfig), ctypes.POINTER(VsgTensor), i32, ctypes.c_char_p,
                                             i32, ctypes.POINTER(ctypes.c_void_p)]
        L.vsg_relenc_workspace_bytes.restype = ctypes.c_size_t
        L.vsg_relenc_workspace_bytes.argtypes = [vp, i32, i32, i32, i32]
        L.vsg_relenc_forward.restype = ctypes.c_int
        L.vsg_relenc_forward.argtypes = [vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, ctypes.c_size_t, vp]
        L.vsg_frame_prior_pack_create.restype = ctypes.c_int
        L.vsg_frame_prior_pack_create.argtypes = L.vsg_relenc_pack_create.argtypes
        L.vsg_frame_prior_workspace_bytes.restype = ctypes.c_size_t
        L.vsg_frame_prior_workspace_bytes.argtypes = [vp, i32, i32, i32]
        L.vsg_frame_prior_forward.restype = ctypes.c_int
        L.vsg_frame_prior_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, ctypes.c_size_t, vp]
        L.vsg_length_regulate.restype = ctypes.c_int
        L.vsg_length_regulate.argtypes = [vp, vp, vp, i32, vp, i32, i32, i32, i32, vp]
        L.vsg_infer_zp.restype = ctypes.c_int
        L.vsg_infer_zp.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, ctypes.c_size_t, vp]
        L.vsg_debug_resblock_bf16.restype = ctypes.c_int
        L.vsg_debug_resblock_bf16.argtypes = [vp, vp, vp, i32, vp, vp, ctypes.c_float, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32]
        L.vsg_debug_set_plan.restype = ctypes.c_int
        L.vsg_debug_set_plan.argtypes = [i32, i32, i32, i32, i32]
        L.vsg_debug_last_ms.restype = ctypes.c_float
        if L.vsg_abi_version() != 1:
            raise RuntimeError("visinger_b200: ABI version mismatch between _lib.py and the shared library")
        _lib = L
        return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vsg_last_error().decode(errors="replace")
        raise RuntimeError(f"visinger_b200: {what} failed ({rc}): {msg}")


def last_launch_count() -> int:
    return int(lib().vsg_last_launch_count())


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"visinger_b200: {name} is on {t.device}; the B200 path has no CPU fallback "
                           "(move the module and its inputs to a CUDA device)")


def as_f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class Pack:
    """Owner of a VsgPack* (device-resident pre-packed weights)."""

    def __init__(self, cfg: VsgConfig, state_dict: Dict[str, torch.Tensor], flow_prefix: str, dec_prefix: str,
                 device: torch.device):
        self._h = ctypes.c_void_p()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("visinger_b200: weights must live on a CUDA device; there is no CPU fallback")
        L = lib()
        host = {}
        for k, v in state_dict.items():
            if (flow_prefix is not None and k.startswith(flow_prefix)) or \
               (dec_prefix is not None and k.startswith(dec_prefix)):
                host[k] = v.detach().to("cpu", torch.float32).contiguous()
        arr = (VsgTensor * len(host))()
        keep = []
        for i, (k, v) in enumerate(host.items()):
            nm = k.encode()
            keep.append(nm)
            arr[i].name = nm
            arr[i].data = v.data_ptr()
            arr[i].ndim = v.dim()
            for d in range(v.dim()):
                arr[i].shape[d] = v.shape[d]
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = L.vsg_pack_create(ctypes.byref(cfg), arr, len(host), (flow_prefix or "").encode(),
                               (dec_prefix or "").encode(), idx, ctypes.byref(self._h))
        check(rc, "vsg_pack_create")
        self.index = idx
        self.hop = int(L.vsg_hop_size(self._h))

    @property
    def handle(self):
        return self._h

    def workspace_bytes(self, B: int, T: int, precision: int) -> int:
        return int(lib().vsg_workspace_bytes(self._h, B, T, precision))

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.vsg_pack_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


def _weight_table(host: Dict[str, torch.Tensor]):
    arr = (VsgTensor * len(host))()
    keep = []
    for i, (k, v) in enumerate(host.items()):
        nm = k.encode()
        keep.append(nm)
        arr[i].name = nm
        arr[i].data = v.data_ptr()
        arr[i].ndim = v.dim()
        for d in range(v.dim()):
            arr[i].shape[d] = v.shape[d]
    return arr, keep


class EncPack:
    """Owner of the VsgPack* of a PosteriorEncoder (vsg_enc_pack_create)."""

    def __init__(self, cfg: VsgEncConfig, state_dict: Dict[str, torch.Tensor], prefix: str, device: torch.device):
        self._h = ctypes.c_void_p()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("visinger_b200: weights must live on a CUDA device; there is no CPU fallback")
        host = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items() if k.startswith(prefix)}
        arr, keep = _weight_table(host)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = lib().vsg_enc_pack_create(ctypes.byref(cfg), arr, len(host), prefix.encode(), idx, ctypes.byref(self._h))
        check(rc, "vsg_enc_pack_create")
        self.index = idx

    @property
    def handle(self):
        return self._h

    def workspace_bytes(self, B: int, T: int, precision: int) -> int:
        return int(lib().vsg_posterior_workspace_bytes(self._h, B, T, precision))

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.vsg_pack_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


class RelEncPack:
    """Owner of the VsgPack* of a RelativeEncoder (vsg_relenc_pack_create) or, with frame_prior=True, of a whole
    FramePriorNetwork (encoder + proj; vsg_frame_prior_pack_create)."""

    def __init__(self, cfg: VsgRelEncConfig, state_dict: Dict[str, torch.Tensor], prefix: str, device: torch.device,
                 frame_prior: bool = False):
        self._h = ctypes.c_void_p()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("visinger_b200: weights must live on a CUDA device; there is no CPU fallback")
        host = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items() if k.startswith(prefix)}
        arr, keep = _weight_table(host)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        create = lib().vsg_frame_prior_pack_create if frame_prior else lib().vsg_relenc_pack_create
        rc = create(ctypes.byref(cfg), arr, len(host), prefix.encode(), idx, ctypes.byref(self._h))
        check(rc, "vsg_frame_prior_pack_create" if frame_prior else "vsg_relenc_pack_create")
        self.index = idx

    @property
    def handle(self):
        return self._h

    def workspace_bytes(self, B: int, T: int, g_per_frame: int, precision: int) -> int:
        return int(lib().vsg_relenc_workspace_bytes(self._h, B, T, g_per_frame, precision))

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.vsg_pack_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


def length_regulate(enc: torch.Tensor, mel2ph: torch.Tensor, pos_table: Optional[torch.Tensor]) -> torch.Tensor:
    """enc [B, H, T_ph] fp32, mel2ph [B, T] int64, pos_table [rows, H] fp32 or None -> [B, H, T]  (vsg_length_regulate)."""
    require_cuda(enc, "enc")
    require_cuda(mel2ph, "mel2ph")
    ec = as_f32c(enc)
    mc = mel2ph.to(torch.int64).contiguous()
    B, H, T_ph = ec.shape
    T = mc.shape[1]
    tab = as_f32c(pos_table) if pos_table is not None else None
    if tab is not None:
        require_cuda(tab, "pos_table")
    y = torch.empty(B, H, T, dtype=torch.float32, device=enc.device)
    if B and T:
        with torch.cuda.device(enc.device):
            rc = lib().vsg_length_regulate(ec.data_ptr(), mc.data_ptr(), tab.data_ptr() if tab is not None else None,
                                           tab.shape[0] if tab is not None else 0, y.data_ptr(), B, H, T_ph, T,
                                           stream_ptr(enc.device))
        check(rc, "vsg_length_regulate")
    return y


_ws_cache: Dict[tuple, torch.Tensor] = {}


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream); allocated through torch's caching allocator."""
    stream = torch.cuda.current_stream(device)
    key = (device.index, stream.cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _ws_cache.pop(key, None)
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def release_workspaces() -> None:
    _ws_cache.clear()


def debug_conv1d_bf16(x_bld: torch.Tensor, w: torch.Tensor, bias, dilation: int, flags: int = 0, add0=None,
                      add1=None, scale: float = 1.0, want_bf16: bool = False, want_f32: bool = True, want_raw: bool = True):
    """Per-layer parity hook: one Conv1d (+ fused epilogue) on the tcgen05 kernel.
    x_bld / add0 / add1: CUDA bf16 [B, L, C] channels-last; w: [Cout, Cin, k] fp32 (any device).
    Returns out_f32 [B, L, Cout], or (out_f32, out_raw_bf16, out_act_bf16) when want_bf16."""
    require_cuda(x_bld, "x")
    assert x_bld.dtype == torch.bfloat16 and x_bld.is_contiguous()
    B, Lx, Cin = x_bld.shape
    planes = 3 if (flags & 256) else 2 if (flags & 4) else 1    # split-bf16: [hi | lo] ([hi | mid | lo]) planes per row
    Cin //= planes
    Cout, _, k = w.shape
    wh = w.detach().to("cpu", torch.float32).contiguous()
    bh = bias.detach().to("cpu", torch.float32).contiguous() if bias is not None else None
    dev = x_bld.device
    out = torch.empty(B, Lx, Cout, dtype=torch.float32, device=dev)
    raw = torch.zeros(B, Lx, planes * Cout, dtype=torch.bfloat16, device=dev) if (want_bf16 and want_raw) else None
    act = torch.zeros(B, Lx, planes * Cout, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    for t in (add0, add1):
        assert t is None or (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous())
    torch.cuda.synchronize(dev)
    rc = lib().vsg_debug_conv1d_bf16(x_bld.data_ptr(), wh.data_ptr(), bh.data_ptr() if bh is not None else None,
                                     add0.data_ptr() if add0 is not None else None,
                                     add1.data_ptr() if add1 is not None else None, float(scale),
                                     out.data_ptr() if want_f32 else None,
                                     raw.data_ptr() if raw is not None else None, act.data_ptr() if want_bf16 else None,
                                     B, Lx, Cin, Cout, k, dilation, flags, dev.index or 0)
    check(rc, "vsg_debug_conv1d_bf16")
    return (out, raw, act) if want_bf16 else out


def debug_pair_bf16(xa_bld: torch.Tensor, w1, b1, w2, b2, d1: int, add0=None, add1=None, scale: float = 1.0,
                    want_f32: bool = True, want_raw: bool = True):
    """Per-layer parity hook of the fused ResBlock1 pair kernel.  xa_bld / add0 / add1: CUDA bf16 [B, L, C];
    returns (out_f32, out_raw_bf16, out_act_bf16)."""
    require_cuda(xa_bld, "xa")
    assert xa_bld.dtype == torch.bfloat16 and xa_bld.is_contiguous()
    B, Lx, C = xa_bld.shape
    k = w1.shape[2]
    h = [t.detach().to("cpu", torch.float32).contiguous() for t in (w1, b1, w2, b2)]
    dev = xa_bld.device
    out = torch.zeros(B, Lx, C, dtype=torch.float32, device=dev) if want_f32 else None
    raw = torch.zeros(B, Lx, C, dtype=torch.bfloat16, device=dev) if want_raw else None
    act = torch.zeros(B, Lx, C, dtype=torch.bfloat16, device=dev)
    torch.cuda.synchronize(dev)
    rc = lib().vsg_debug_pair_bf16(xa_bld.data_ptr(), h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(),
                                   add0.data_ptr() if add0 is not None else None,
                                   add1.data_ptr() if add1 is not None else None, float(scale),
                                   out.data_ptr() if want_f32 else None,
                                   raw.data_ptr() if raw is not None else None, act.data_ptr(), B, Lx, C, k, d1,
                                   dev.index or 0)
    check(rc, "vsg_debug_pair_bf16")
    return out, raw, act


def debug_resblock_bf16(xa_bld: torch.Tensor, ws, bs, dilations, add1=None, scale: float = 1.0, max_mb: int = 0, sets: int = 0,
                        want_raw: bool = True, want_act: bool = True, reps: int = 1, want_f32: bool = True):
    """Per-layer parity hook of the whole-ResBlock1 kernel (csrc/rb_tc.cuh).  xa_bld / add1: CUDA bf16 [B, L, C];
    ws / bs: lists [c1_0, c2_0, c1_1, c2_1, ...] of fp32 [C, C, k] / [C].  Returns (out_f32, out_raw_bf16, out_act_bf16[, ms]).
    sets bit 11 (with bit 8): the row-packed kernel's split-bf16 instantiation -- xa / add1 / raw / act are two-plane
    tensors [B, L, 2 C] (`split_bf16`), out_f32 stays [B, L, C].  (The kernel takes xa / add1 and writes raw as planar
    tensors [2][B, L, C]; the conversion from / to the decoder's rows [hi | lo] happens here.)"""
    require_cuda(xa_bld, "xa")
    assert xa_bld.dtype == torch.bfloat16 and xa_bld.is_contiguous()
    B, Lx, C = xa_bld.shape
    x3 = bool(sets & 2048)
    if x3:
        assert C % 2 == 0
        C //= 2
        xa_bld = torch.stack([xa_bld[..., :C], xa_bld[..., C:]]).contiguous()
        if add1 is not None:
            add1 = torch.stack([add1[..., :C], add1[..., C:]]).contiguous()
    k = ws[0].shape[2]
    n_pairs = len(dilations)
    assert len(ws) == 2 * n_pairs and len(bs) == 2 * n_pairs
    wh = torch.stack([w.detach().to("cpu", torch.float32) for w in ws]).contiguous()
    bh = torch.stack([b.detach().to("cpu", torch.float32) for b in bs]).contiguous()
    dl = (ctypes.c_int32 * n_pairs)(*[int(d) for d in dilations])
    dev = xa_bld.device
    out = torch.zeros(B, Lx, C, dtype=torch.float32, device=dev) if want_f32 else None
    raw = torch.zeros((2, B, Lx, C) if x3 else (B, Lx, C), dtype=torch.bfloat16, device=dev) if want_raw else None
    act = torch.zeros(B, Lx, (2 if x3 else 1) * C, dtype=torch.bfloat16, device=dev) if want_act else None
    assert add1 is None or (add1.is_cuda and add1.dtype == torch.bfloat16 and add1.is_contiguous())
    torch.cuda.synchronize(dev)
    if reps > 1:
        lib().vsg_debug_set_plan(0, 0, -1, -1, reps)
    try:
        rc = lib().vsg_debug_resblock_bf16(xa_bld.data_ptr(), wh.data_ptr(), bh.data_ptr(), n_pairs, dl,
                                           add1.data_ptr() if add1 is not None else None, float(scale),
                                           out.data_ptr() if out is not None else None, raw.data_ptr() if raw is not None else None,
                                           act.data_ptr() if act is not None else None, B, Lx, C, k, int(max_mb), int(sets),
                                           dev.index or 0)
        ms = float(lib().vsg_debug_last_ms()) if reps > 1 else None
    finally:
        if reps > 1:
            lib().vsg_debug_set_plan(0, 0, -1, -1, 1)
    check(rc, "vsg_debug_resblock_bf16")
    if x3 and raw is not None:
        raw = torch.cat([raw[0], raw[1]], dim=-1)
    return (out, raw, act, ms) if reps > 1 else (out, raw, act)


def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """[..., C] fp32 -> [..., 2C] bf16 planes [hi | lo] with hi = bf16(x), lo = bf16(x - hi)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=-1).contiguous()


def merge_bf16(x: torch.Tensor) -> torch.Tensor:
    """Inverse view of split_bf16: [..., 2C] bf16 planes -> [..., C] fp32 = hi + lo."""
    C = x.shape[-1] // 2
    return x[..., :C].float() + x[..., C:].float()


def set_tc_options(halo_mode: int = 1, w_resident: int = 1, l2_tensor_mb: int = -1, min_tiles: int = -1) -> None:
    check(lib().vsg_set_tc_options(int(halo_mode), int(w_resident), int(l2_tensor_mb), int(min_tiles)),
          "vsg_set_tc_options")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
