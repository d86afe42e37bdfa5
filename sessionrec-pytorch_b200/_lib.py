"""ctypes binding of libsessrec_b200.so.  Signatures are parsed from include/sessrec_b200.h, so the header is
the single source of truth for the C ABI.  There is NO fallback: a missing library raises."""
import ctypes
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / 'include' / 'sessrec_b200.h'
LIB_PATH = PKG / 'libsessrec_b200.so'

HEADS = 8
MAX_GAT_INST = 8
NORM_NONE, NORM_NISER, NORM_L2, NORM_EPS = 0, 1, 2, 3
SITE_EMBED, SITE_READOUT, SITE_GGNN = 0x100, 0x200, 0x300
SITE_GAT_SRC, SITE_GAT_DST, SITE_GAT_ATTN = 0x1000, 0x1001, 0x1002


class Dropout(ctypes.Structure):
    _fields_ = [('p', ctypes.c_float), ('site', ctypes.c_uint32), ('seed', ctypes.c_uint64)]


class GatInst(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ('in_ptr', 'in_src', 'in_eid', 'out_ptr', 'out_dst', 'out_eid', 'Zel', 'er', 'bias', 'xdst', 'att',
                 'dedge', 'der', 'dZel')] + \
               [('n_src', ctypes.c_int), ('n_dst', ctypes.c_int), ('n_edges', ctypes.c_int), ('attn_site', ctypes.c_uint32)]


_SCALARS = {'int': ctypes.c_int, 'long long': ctypes.c_longlong, 'float': ctypes.c_float, 'uint32_t': ctypes.c_uint32,
            'uint64_t': ctypes.c_uint64}


def parse_header(path=HEADER):
    """[(name, restype, [(ctype, argname), ...])] for every function declared in the header."""
    text = re.sub(r'/\*.*?\*/', ' ', Path(path).read_text(), flags=re.S)
    text = re.sub(r'#.*', ' ', text)
    text = re.sub(r'typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;', ' ', text, flags=re.S)
    out = []
    for m in re.finditer(r'([\w\s\*]+?)\b(srk_\w+)\s*\(([^;{}]*?)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                mm = re.match(r'(.*?)(\w+)$', a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        out.append((name, ret, parsed))
    return out


def _ctype(t):
    t = t.replace('const ', '').strip()
    if t.endswith('*'):
        return ctypes.c_char_p if t == 'char*' else ctypes.c_void_p
    return _SCALARS[t]


class SessRecError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not LIB_PATH.exists():
            raise SessRecError(
                f'{LIB_PATH} is missing: build it with `python sessionrec-pytorch_b200/build.py` '
                '(or __graft_entry__.build()). There is no CPU / PyTorch fallback for this path.')
        self._dll = ctypes.CDLL(str(LIB_PATH))
        self.functions = {}
        for name, ret, args in parse_header():
            fn = getattr(self._dll, name)           # AttributeError if the header declares an unexported symbol
            fn.argtypes = [_ctype(t) for t, _ in args]
            fn.restype = ctypes.c_char_p if ret.replace('const ', '').strip() == 'char*' else _SCALARS[ret.strip()]
            self.functions[name] = fn

    def last_error(self):
        return (self.functions['srk_last_error']() or b'').decode()

    def call(self, name, *args):
        r = self.functions[name](*args)
        if r < 0:
            raise SessRecError(f'{name} failed ({r}): {self.last_error()}')
        return r


_LIB = None
LAUNCHES = 0     # number of C-ABI compute calls issued (each enqueues >= 1 of our kernels); bench reports kernels


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def byref(s):
    return None if s is None else ctypes.cast(ctypes.pointer(s), ctypes.c_void_p)
