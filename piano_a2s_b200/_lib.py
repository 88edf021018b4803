"""ctypes binding of libpa2s.so.  Signatures are taken from include/pa2s.h so header, library and binding cannot drift.

There is NO fallback: if the library is missing or a call returns non-zero this raises."""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PA2S_LIB") or os.path.join(HERE, "libpa2s.so")      # PA2S_LIB: an alternative build (kernel experiments)
HEADER = os.path.join(os.path.dirname(HERE), "include", "pa2s.h")

_SCALARS = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
            "unsigned long long": ctypes.c_ulonglong}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every PA2S_API prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"PA2S_API\s+([\w\s]+?)\s*(\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = " ".join(a.split()[:-1]).replace("const ", "").strip()
                    argtypes.append(_SCALARS[ty])
        protos[name] = (_SCALARS[ret], argtypes)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m piano_a2s_b200.build` "
                                   "(there is no CPU / PyTorch fallback for the hot path)")
            dll = ctypes.CDLL(LIB_PATH)
            for name, (ret, argtypes) in parse_header().items():
                fn = getattr(dll, name)
                fn.restype = ret
                fn.argtypes = argtypes
            self._dll = dll
        return self._dll

    def __getattr__(self, name):
        fn = getattr(self.load(), name)
        if name in ("pa2s_launch_count", "pa2s_dec_args_size", "pa2s_gru_seq_max_bg", "pa2s_conv3x3_num_partials",
                    "pa2s_tc_conv_pack_bytes", "pa2s_tc_conv_num_partials", "pa2s_gemm_tc_supported",
                    "pa2s_tc_conv_wgrad_num_partials", "pa2s_dec_persist_grid", "pa2s_dec_deferred_blocks", "pa2s_planes_bytes",
                    "pa2s_conv_tma_num_partials", "pa2s_conv_tma_wgrad_num_partials", "pa2s_decm_args_size", "pa2s_decm_max_queries",
                    "pa2s_decm_tile_max", "pa2s_decm_grid", "pa2s_decm_deferred_blocks", "pa2s_conv_tma_get_impl"):
            return fn

        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise RuntimeError(f"{name} failed with code {rc}")
        return call


lib = _Lib()


_DEVICE_INDEX = None


def stream():
    """Raw handle of torch's CURRENT stream on this process's device (one process per GPU).  Goes through the C binding
    directly: torch.cuda.current_stream() costs ~5 us of Python per call, which at ~1500 kernel launches per training step
    was 8 ms of host time per step."""
    global _DEVICE_INDEX
    if _DEVICE_INDEX is None:
        _DEVICE_INDEX = torch.cuda.current_device()
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(_DEVICE_INDEX))


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "libpa2s only takes device tensors"
    return ctypes.c_void_p(t.data_ptr())


class DecArgs(ctypes.Structure):
    """Mirror of `struct DecArgs` in csrc/dec_args.cuh (field order and types must match exactly)."""
    _ints = ["B", "T", "V", "VP", "S", "max_steps", "NS", "tile", "inference", "save", "tile_pad", "reserved_"]
    _ptrs = ["enc", "Ep", "Wattn", "v", "emb", "W_ih", "W_hh", "b_ih", "b_hh", "W_out", "b_out",
             "W_outT", "W_hT", "W_ihT", "W_hhT", "gt", "use_gt", "mask", "logp", "lengths", "eos", "counters",
             "hs", "ctxs", "attn", "gates", "qs", "xtok", "toks", "xbuf", "hc", "logits", "pm", "pl", "pc", "tickets",
             "dlogp", "dlogits_all", "dgi_all", "dgh_all", "dq_all", "dctx_all", "dxtok_all", "dEp", "dv_part",
             "d_hc", "dhq", "dx", "dq_part", "dh_carry", "dh_last", "sync", "dhc_all", "ds_all", "dv", "prof"]
    _fields_ = [(n, ctypes.c_int) for n in _ints] + [(n, ctypes.c_void_p) for n in _ptrs]


def make_dec_args(**kw):
    a = DecArgs()
    for k, v in kw.items():
        if k in DecArgs._ints:
            setattr(a, k, int(v))
        else:
            assert k in DecArgs._ptrs, k
            setattr(a, k, None if v is None else v.data_ptr())
    return a


class DecMArgs(ctypes.Structure):
    """Mirror of `struct DecMArgs` in csrc/decm_args.cuh (field order and types must match exactly)."""
    _ints = ["B", "NQ", "T", "V", "VP", "S", "max_steps", "NS", "tile", "inference", "save", "Rtot", "r0", "bars", "k0", "Spitch", "tc"]
    _ptrs = ["enc", "Ee", "Wattn", "v", "emb", "W_ih", "W_hh", "b_ih", "b_hh", "W_out", "b_out", "W_hT", "W_ihT", "W_hhT",
             "gt", "mask", "logp", "lengths", "eos", "counters",
             "hs", "ctxs", "attn", "gates", "qs", "eqs", "xtok", "toks", "ml",
             "xbuf", "logits", "pm", "pl", "pc", "tickets", "sync",
             "dhc_all", "dgi_all", "dgh_all", "dq_all", "dctx_all", "dxtok_all", "ds_all", "dEp", "dv_part", "d_hc", "dhq", "dx",
             "dq_part", "dh_carry", "dlogp", "dlogits_all", "prof"]
    _fields_ = [(n, ctypes.c_int) for n in _ints] + [("Sq", ctypes.c_int * 8), ("tf_bits", ctypes.c_uint * 64), ("has_tf", ctypes.c_int),
                                                     ("pad_", ctypes.c_int)] + [(n, ctypes.c_void_p) for n in _ptrs]


def make_decm_args(Sq, tf_bits=None, **kw):
    """Pointer fields take a tensor, a raw int address (pre-offset views) or None.  tf_bits: python int bit mask (bit q * Spitch + s) or None."""
    a = DecMArgs()
    assert len(Sq) <= 8
    for i, v in enumerate(Sq):
        a.Sq[i] = int(v)
    if tf_bits is not None:
        assert tf_bits >> 2048 == 0, "teacher-forcing mask of a launch is limited to 2048 (query, step) pairs"
        a.has_tf = 1
        for i in range(64):
            a.tf_bits[i] = (tf_bits >> (32 * i)) & 0xFFFFFFFF
    for k, v in kw.items():
        if k in DecMArgs._ints:
            setattr(a, k, int(v))
        else:
            assert k in DecMArgs._ptrs, k
            setattr(a, k, None if v is None else (v if isinstance(v, int) else v.data_ptr()))
    return a
