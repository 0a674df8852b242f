"""ctypes binding of ``libtitanet_sm100.so`` (the C ABI in ``include/titanet_b200.h``).

The argument types of every entry point are read from the header itself, so the
binding cannot drift from the declaration.  There is no fallback: if the library is
missing the first op raises (``TitanetLibraryError``), and every op refuses CPU
tensors.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libtitanet_sm100.so")
HEADER_CANDIDATES = (os.path.join(ROOT, "include", "titanet_b200.h"), os.path.join(HERE, "titanet_b200.h"))


class TitanetLibraryError(RuntimeError):
    pass


_CTYPE = {
    "int": ctypes.c_int, "unsigned int": ctypes.c_uint, "float": ctypes.c_float, "double": ctypes.c_double,
    "long long": ctypes.c_longlong, "unsigned long long": ctypes.c_ulonglong, "size_t": ctypes.c_size_t,
}


def header_path() -> str:
    for p in HEADER_CANDIDATES:
        if os.path.exists(p):
            return p
    raise TitanetLibraryError("titanet_b200.h not found next to the package")


def parse_header(path: str | None = None) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """name -> (return type, [(ctype string, arg name), ...]) for every prototype."""
    text = open(path or header_path()).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(tn_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                alist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, alist)
    return protos


def _to_ctype(t: str):
    if "*" in t:
        return ctypes.c_void_p
    t = t.replace("const ", "").strip()
    return _CTYPE[t]


class _Lib:
    def __init__(self):
        self._dll = None
        self._fns = {}

    def load(self):
        if self._dll is not None:
            return self
        if not os.path.exists(LIB_PATH):
            raise TitanetLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m titanet_b200._build` "
                "(needs nvcc; there is no CPU or PyTorch fallback)")
        self._dll = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in parse_header().items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError as e:  # pragma: no cover
                raise TitanetLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = (ctypes.c_char_p if ret.startswith("const char") else
                          ctypes.c_longlong if ret == "long long" else ctypes.c_int)
            fn.argtypes = [_to_ctype(t) for t, _ in args]
            self._fns[name] = fn
        return self

    def last_error(self) -> str:
        msg = self._fns["tn_last_error"]()
        return msg.decode() if msg else ""

    def call(self, name: str, *args):
        fn = self._fns.get(name)
        if fn is None:
            self.load()
            fn = self._fns[name]
        rc = fn(*args)
        if rc != 0:
            raise TitanetLibraryError(f"{name} failed (code {rc}): {self.last_error()}")

    def query(self, name: str, *args) -> int:
        """Entry points that return a size (``long long``) instead of a status."""
        fn = self._fns.get(name)
        if fn is None:
            self.load()
            fn = self._fns[name]
        return int(fn(*args))


class TnBnFold(ctypes.Structure):
    """``tn_bn_fold`` of include/titanet_b200.h (train-mode BatchNorm folded by its producer kernel)."""
    _fields_ = [("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("running_mean", ctypes.c_void_p),
                ("running_var", ctypes.c_void_p), ("num_batches_tracked", ctypes.c_void_p),
                ("momentum", ctypes.c_float), ("eps", ctypes.c_float), ("n", ctypes.c_double),
                ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p), ("mean", ctypes.c_void_p),
                ("invstd", ctypes.c_void_p)]


class TnBnBwd(ctypes.Structure):
    """``tn_bn_bwd`` of include/titanet_b200.h (BatchNorm backward folded into the data-gradient GEMM's operand load)."""
    _fields_ = [("z", ctypes.c_void_p), ("dscale", ctypes.c_void_p), ("dshift", ctypes.c_void_p), ("mean", ctypes.c_void_p),
                ("invstd", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("n", ctypes.c_double), ("g_out", ctypes.c_void_p),
                ("dbias", ctypes.c_void_p), ("dgamma", ctypes.c_void_p), ("dbeta", ctypes.c_void_p)]


class TnScratch(ctypes.Structure):
    """``tn_scratch`` of include/titanet_b200.h (workspace + tickets of the atomics-free cross-block reductions)."""
    _fields_ = [("parts", ctypes.c_void_p), ("parts_floats", ctypes.c_longlong), ("tickets", ctypes.c_void_p),
                ("accum", ctypes.c_void_p), ("accum_words", ctypes.c_longlong)]


class TnSplitJob(ctypes.Structure):
    """``tn_split_job`` of include/titanet_b200.h."""
    _fields_ = [("W", ctypes.c_void_p), ("ws", ctypes.c_void_p), ("M", ctypes.c_int), ("Kd", ctypes.c_int),
                ("transpose", ctypes.c_int), ("tile0", ctypes.c_int)]


class TnAdamJob(ctypes.Structure):
    """``tn_adam_job`` of include/titanet_b200.h."""
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("n", ctypes.c_longlong)]


LIB = _Lib()
_checked_device = False


def require_cuda(*tensors: torch.Tensor):
    """No CPU path exists: reject anything that is not a CUDA tensor."""
    global _checked_device
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TitanetLibraryError(
                "titanet_b200 runs on NVIDIA B200 (sm_100a) only: got a CPU tensor and there is no CPU fallback")
    if not _checked_device:
        LIB.load()
        LIB.call("tn_device_check")
        _checked_device = True


def ptr(t):
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# launch accounting (bench.py reads these): ABI calls and, optionally, per-call device time
COUNTS: Dict[str, int] = {}
KERNELS_PER_CALL = {"tn_zero": 0, "tn_ce_fwd_bwd": 2, "tn_margin_fwd_bwd": 2}
_profile = None          # list of (key, start_event, stop_event) while profiling


def profile_start():
    global _profile
    _profile = []


def profile_stop() -> Dict[str, Tuple[int, float]]:
    """key -> (launches, total milliseconds); synchronises the device."""
    global _profile
    rec, _profile = _profile or [], None
    torch.cuda.synchronize()
    out: Dict[str, Tuple[int, float]] = {}
    for key, e0, e1 in rec:
        n, ms = out.get(key, (0, 0.0))
        out[key] = (n + 1, ms + e0.elapsed_time(e1))
    return out


def kernel_launches() -> int:
    return sum(n * KERNELS_PER_CALL.get(k, 1) for k, n in COUNTS.items())


def call(name: str, *args, tag: str = ""):
    COUNTS[name] = COUNTS.get(name, 0) + 1
    if _profile is None:
        LIB.call(name, *args, stream())
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    LIB.call(name, *args, stream())
    e1.record()
    _profile.append((f"{name}[{tag}]" if tag else name, e0, e1))
