"""ctypes binding of the C ABI declared in include/dsgan_b200.h.

Prototypes are parsed from the header itself, so the Python side cannot drift from the ABI.  There is no
fallback of any kind: a missing library or a non-sm_100 device raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "dsgan_b200.h")
LIB_PATH = os.path.join(HERE, "libdsgan_b200.so")


class ConvDesc(ctypes.Structure):
    """Mirror of dsgan_conv_desc."""
    _fields_ = [(n, ctypes.c_int) for n in
                ("dtype", "N", "Hi", "Wi", "Ci", "Ho", "Wo", "Co", "kh", "kw", "stride", "pad", "transposed",
                 "ld_in", "ld_out", "ld_aux", "ld_pre")] + \
               [(n, ctypes.c_longlong) for n in ("w_sco", "w_sci", "w_sky", "w_skx")] + \
               [(n, ctypes.c_int) for n in ("act", "dact", "accumulate")]


class TcConvDesc(ctypes.Structure):
    """Mirror of dsgan_tc_conv_desc."""
    _fields_ = [(n, ctypes.c_int) for n in
                ("N", "Hi", "Wi", "Ci", "ld_in", "Ho", "Wo", "Co", "ld_out", "Hg", "Wg", "in_stride", "out_stride",
                 "nclass")] + \
               [("oy0", ctypes.c_int * 4), ("ox0", ctypes.c_int * 4), ("ntaps", ctypes.c_int * 4)] + \
               [(n, ctypes.c_int) for n in ("nslabs", "co_pad", "ci_pad")] + \
               [("dy", ctypes.c_int * 64), ("dx", ctypes.c_int * 64), ("slab", ctypes.c_int * 64)] + \
               [(n, ctypes.c_int) for n in ("ld_aux", "ld_pre", "act", "dact", "accumulate")]


class PackJob(ctypes.Structure):
    """Mirror of dsgan_pack_job."""
    _fields_ = [("src", ctypes.c_ulonglong), ("dst", ctypes.c_ulonglong)] + \
               [(n, ctypes.c_int) for n in ("O", "I", "O_pad", "I_pad", "kh", "kw")] + \
               [(n, ctypes.c_longlong) for n in ("s_o", "s_i", "s_ky", "s_kx")] + \
               [("block0", ctypes.c_int), ("pad_", ctypes.c_int)]


class TcWgradDesc(ctypes.Structure):
    """Mirror of dsgan_tc_wgrad_desc."""
    _fields_ = [(n, ctypes.c_int) for n in ("N", "Hg", "Wg", "Cg", "ld_g", "Hx", "Wx", "Cx", "ld_x", "x_stride", "ntaps")] + \
               [("dy", ctypes.c_int * 16), ("dx", ctypes.c_int * 16), ("tap_off", ctypes.c_longlong * 16),
                ("s_g", ctypes.c_longlong), ("s_x", ctypes.c_longlong)]


class DwBranch(ctypes.Structure):
    """Mirror of dsgan_dw_branch."""
    _fields_ = [("w", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("dw", ctypes.c_void_p), ("db", ctypes.c_void_p),
                ("k", ctypes.c_int), ("c0", ctypes.c_int), ("c", ctypes.c_int), ("pad_", ctypes.c_int)]


_SCALARS = {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
            "unsigned long long": ctypes.c_ulonglong, "size_t": ctypes.c_size_t, "unsigned": ctypes.c_uint}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every `dsgan_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(dsgan_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret or "struct" in ret:
            continue
        restype = ctypes.c_char_p if "char" in ret else _SCALARS.get(ret.replace("const ", "").strip(), ctypes.c_int)
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = " ".join(a.replace("const ", "").split()[:-1])
                    argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


class DsganError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise DsganError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(dsgan_b200 has no CPU or library fallback)" % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = restype, argtypes
        if self.cdll.dsgan_abi_version() != 1:
            raise DsganError("ABI version mismatch")
        self.profiler = None  # optional engine.Profile: CUDA events around every ABI call

    def last_error(self):
        return self.cdll.dsgan_last_error().decode()

    def __getattr__(self, short):
        """lib.conv_fwd(...) -> dsgan_conv_fwd(...) with error checking."""
        fn = getattr(self.cdll, "dsgan_" + short)
        if fn.restype is not ctypes.c_int:
            return fn

        def call(*args):
            prof = self.profiler
            if prof is not None:
                tok = prof.begin(short)
            if fn(*args) != 0:
                raise DsganError("dsgan_%s: %s" % (short, self.last_error()))
            if prof is not None:
                prof.end(tok)
        call.__name__ = short
        setattr(self, short, call)
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def require_device():
    """Fail loudly unless a B200 is the current device."""
    import torch
    if not torch.cuda.is_available():
        raise DsganError("dsgan_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib().device_check()
