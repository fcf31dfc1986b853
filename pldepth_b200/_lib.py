"""ctypes binding of libpldepth_b200.so (the C ABI declared in include/pldepth_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA device is not
available, the product path raises.  Build it with ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C pldepth_b200/csrc``.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLDEPTH_B200_LIB") or os.path.join(_HERE, "libpldepth_b200.so")

PLD_MAX_RANKING_SIZE = 512
PLD_MAX_PIXELS = 1 << 23
ST_EMPTY_MASK, ST_BAD_INDEX, ST_MT_EXHAUSTED, ST_INTERNAL = 1, 2, 4, 8
STRATEGY = {"purely": 0, "masked": 1, "thresholded": 2, "information": 3}
PROMOTION = {"nep50": 0, "legacy": 1}

c_void_p, c_int, c_float, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double
c_u64, c_i64, c_u32 = ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32

# name -> (restype, argtypes); every symbol include/pldepth_b200.h declares
SIGNATURES = {
    "pld_version": (c_int, []),
    "pld_last_error": (ctypes.c_char_p, []),
    "pld_launch_count": (c_u64, []),
    "pld_ctx_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "pld_ctx_destroy": (c_int, [c_void_p]),
    "pld_ctx_status": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_int)]),
    "pld_ctx_set_deterministic": (c_int, [c_void_p, c_int]),
    "pld_ctx_device_offset": (c_int, [c_void_p, c_int, c_u64]),
    "pld_ctx_kernel_timing": (c_int, [c_void_p, c_int]),
    "pld_ctx_kernel_times": (c_int, [c_void_p, ctypes.POINTER(c_float), c_int, ctypes.POINTER(c_int)]),
    "pld_mask_compact": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p]),
    "pld_sample_lists_philox": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_int, c_u64, c_u64, c_int, c_void_p, c_void_p, c_void_p]),
    "pld_sample_lists_fed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    "pld_sample_lists_mt": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pld_mt19937_init": (c_int, [c_void_p, c_u32, c_void_p, c_void_p, c_void_p]),
    "pld_mt19937_generate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "pld_gt_minmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pld_score_lists": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_double,
                                c_int, c_void_p, c_void_p]),
    "pld_select_top": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p]),
    "pld_listmle_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pld_fused_sample_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                          c_int, c_int, c_int, c_u64, c_u64, c_int, c_float, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pld_fused_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_u64, c_u64, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "pld_fused_step_m8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_u64, c_u64, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "pld_fused_step_scored": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_double, c_double, c_int, c_u64, c_u64, c_int, c_float,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "pld_gather_predictions": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p]),
    "pld_ordinal_error": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                  c_void_p]),
    "pld_eval_ordinal_pairs_mt": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int, c_int, c_void_p,
                                          c_i64, c_void_p, c_void_p, c_void_p]),
    "pld_eval_invert_rankings": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "pld_ndcg": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}


class PLDError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()


def load_library(path=None):
    """dlopen the C-ABI library and attach signatures.  Raises if it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.isfile(p):
            raise PLDError(
                "pldepth_b200: %s not found -- the CUDA library is not built and there is no CPU "
                "fallback.  Run `python -c \"import __graft_entry__ as g; g.build()\"`." % p)
        lib = ctypes.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load_library().pld_last_error()
        raise PLDError("pldepth_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


class Context(object):
    """One pld_ctx (device scratch + status word).  Not shareable across threads: use
    ``Context.current(device)`` which keeps one per (thread, device)."""

    _tls = threading.local()

    def __init__(self, device):
        self.lib = load_library()
        self.device = int(device)
        h = c_void_p()
        check(self.lib.pld_ctx_create(self.device, ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pld_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @classmethod
    def current(cls, device):
        d = getattr(cls._tls, "ctxs", None)
        if d is None:
            d = cls._tls.ctxs = {}
        ctx = d.get(int(device))
        if ctx is None:
            ctx = d[int(device)] = cls(device)
        return ctx

    def status(self, stream_ptr):
        """Synchronise the stream and return (and clear) the PLD_ST_* bits."""
        s = c_int(0)
        check(self.lib.pld_ctx_status(self.handle, c_void_p(stream_ptr), ctypes.byref(s)))
        return s.value

    def set_deterministic(self, on=True):
        """Bit-reproducible gradients (64-bit fixed-point accumulation) for this context."""
        check(self.lib.pld_ctx_set_deterministic(self.handle, 1 if on else 0))

    def device_offset(self, enable=True, start=0):
        """Keep the Philox offset in device memory (advanced by the library after every step): makes the
        step capturable in a CUDA graph."""
        check(self.lib.pld_ctx_device_offset(self.handle, 1 if enable else 0, int(start)))

    def kernel_timing(self, slots):
        check(self.lib.pld_ctx_kernel_timing(self.handle, int(slots)))

    def kernel_times(self, capacity=4096):
        buf = (c_float * capacity)()
        n = c_int(0)
        check(self.lib.pld_ctx_kernel_times(self.handle, buf, capacity, ctypes.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def raise_on_status(self, stream_ptr):
        s = self.status(stream_ptr)
        if s & ST_EMPTY_MASK:
            # the reference raises ValueError from np.random.randint(0) (sampling.py:113)
            raise ValueError("low >= high: an image has no valid mask pixel")
        if s & ST_BAD_INDEX:
            raise IndexError("ranking index outside the prediction map / valid-pixel table")
        if s & ST_MT_EXHAUSTED:
            raise PLDError("MT19937 word stream exhausted")
        if s & ST_INTERNAL:
            raise PLDError("internal self-check failed (top-R selection); the results of the call are void")
        return s


def launch_count():
    return int(load_library().pld_launch_count())
