"""fluent-bit_b200 -- Python mirror (ctypes) of the libflbgpu C ABI (include/flbgpu.h).

The product is the C-ABI library `libflbgpu.so` next to this file (CUDA, sm_100a).  This
module only binds it for tests and bench.py; it contains no compute and no fallback:
importing works without a GPU, but `Context()` raises when the library is missing or no
CUDA device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libflbgpu.so")

FILTER_MODIFIED = 1
FILTER_NOTOUCH = 2
TYPE_INT, TYPE_FLOAT, TYPE_BOOL, TYPE_STRING, TYPE_HEX = 1, 2, 3, 4, 5
_TYPE_NAMES = {"integer": TYPE_INT, "float": TYPE_FLOAT, "bool": TYPE_BOOL, "string": TYPE_STRING, "hex": TYPE_HEX}


class FlbGpuError(RuntimeError):
    pass


class ParserTypes(C.Structure):
    _fields_ = [("key", C.c_char_p), ("key_len", C.c_int), ("type", C.c_int)]


class Time(C.Structure):
    _fields_ = [("tv_sec", C.c_int64), ("tv_nsec", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("records_in", C.c_uint64), ("records_out", C.c_uint64), ("bytes_in", C.c_uint64),
                ("bytes_out", C.c_uint64), ("kernel_launches", C.c_uint64), ("passes", C.c_uint32),
                ("error_bits", C.c_uint32)]


EXPORTS = [
    "flbgpu_init", "flbgpu_shutdown", "flbgpu_last_error", "flbgpu_backend_name", "flbgpu_device_count",
    "flbgpu_parser_create", "flbgpu_parser_get", "flbgpu_parser_do", "flbgpu_parser_destroy",
    "flbgpu_filter_new", "flbgpu_filter_set_property", "flbgpu_filter_init", "flbgpu_filter_cb",
    "flbgpu_filter_destroy", "flbgpu_chain_new", "flbgpu_chain_add", "flbgpu_chain_init", "flbgpu_chain_do",
    "flbgpu_chain_destroy", "flbgpu_chain_do_device", "flbgpu_chain_stats", "flbgpu_dev_alloc",
    "flbgpu_dev_free", "flbgpu_dev_upload", "flbgpu_dev_download", "flbgpu_host_alloc", "flbgpu_host_free",
    "flbgpu_stream", "flbgpu_kernel_ms",
]


def load(path=None):
    """dlopen the C-ABI library and declare prototypes."""
    path = path or PRODUCT_LIB
    if not os.path.exists(path):
        raise FlbGpuError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
    L.flbgpu_init.restype = vp; L.flbgpu_init.argtypes = [C.c_int]
    L.flbgpu_shutdown.argtypes = [vp]
    L.flbgpu_last_error.restype = cp
    L.flbgpu_backend_name.restype = cp
    L.flbgpu_parser_create.restype = vp
    L.flbgpu_parser_create.argtypes = [vp, cp, cp, cp, C.c_int, cp, cp, cp, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(ParserTypes), C.c_int, vp]
    L.flbgpu_parser_get.restype = vp; L.flbgpu_parser_get.argtypes = [vp, cp]
    L.flbgpu_parser_do.argtypes = [vp, cp, sz, C.POINTER(vp), C.POINTER(sz), C.POINTER(Time)]
    L.flbgpu_parser_destroy.argtypes = [vp]
    L.flbgpu_filter_new.restype = vp; L.flbgpu_filter_new.argtypes = [vp, cp]
    L.flbgpu_filter_set_property.argtypes = [vp, cp, cp]
    L.flbgpu_filter_init.argtypes = [vp]
    L.flbgpu_filter_cb.argtypes = [vp, vp, sz, cp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    L.flbgpu_filter_destroy.argtypes = [vp]
    L.flbgpu_chain_new.restype = vp; L.flbgpu_chain_new.argtypes = [vp]
    L.flbgpu_chain_add.argtypes = [vp, vp]
    L.flbgpu_chain_init.argtypes = [vp]
    L.flbgpu_chain_do.argtypes = [vp, vp, sz, cp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    L.flbgpu_chain_destroy.argtypes = [vp]
    L.flbgpu_chain_do_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.flbgpu_chain_stats.argtypes = [vp, C.POINTER(Stats)]
    L.flbgpu_dev_alloc.restype = vp; L.flbgpu_dev_alloc.argtypes = [vp, sz]
    L.flbgpu_dev_free.argtypes = [vp, vp]
    L.flbgpu_dev_upload.argtypes = [vp, vp, vp, sz]
    L.flbgpu_dev_download.argtypes = [vp, vp, vp, sz]
    L.flbgpu_host_alloc.restype = vp; L.flbgpu_host_alloc.argtypes = [vp, sz]
    L.flbgpu_host_free.argtypes = [vp, vp]
    L.flbgpu_stream.restype = vp; L.flbgpu_stream.argtypes = [vp]
    L.flbgpu_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _b(s):
    if s is None:
        return None
    return s if isinstance(s, bytes) else s.encode()


class Context:
    """flbgpu_ctx: parsers + filters of one device (mirrors struct flb_config for this path)."""

    def __init__(self, device=0, lib=None):
        self.L = lib if lib is not None else load()
        self.h = self.L.flbgpu_init(device)
        if not self.h:
            raise FlbGpuError("flbgpu_init(%d) failed: %s" % (device, self.err()))

    def err(self):
        e = self.L.flbgpu_last_error()
        return e.decode(errors="replace") if e else ""

    def parser(self, name, format, regex=None, skip_empty=True, time_fmt=None, time_key=None, time_offset=None,
               time_keep=False, time_strict=True, logfmt_no_bare_keys=False, types=None):
        """flb_parser_create(); `types` is the Types option text, e.g. "code:integer size:integer"."""
        arr, n = None, 0
        if types:
            items = [t.split(":") for t in types.split()]
            arr = (ParserTypes * len(items))()
            for i, (k, t) in enumerate(items):
                arr[i].key = _b(k); arr[i].key_len = len(_b(k)); arr[i].type = _TYPE_NAMES.get(t.lower(), TYPE_STRING)
            n = len(items)
        p = self.L.flbgpu_parser_create(self.h, _b(name), _b(format), _b(regex), int(skip_empty), _b(time_fmt),
                                        _b(time_key), _b(time_offset), int(time_keep), int(time_strict), 0,
                                        int(logfmt_no_bare_keys), arr, n, None)
        if not p:
            raise FlbGpuError("parser_create(%s): %s" % (name, self.err()))
        return Parser(self, p)

    def filter(self, plugin, props):
        """flb_filter_new + set_property (props: ordered list of (key, value)) + cb_init."""
        f = self.L.flbgpu_filter_new(self.h, _b(plugin))
        if not f:
            raise FlbGpuError("filter_new(%s): %s" % (plugin, self.err()))
        for k, v in props:
            self.L.flbgpu_filter_set_property(f, _b(k), _b(v))
        if self.L.flbgpu_filter_init(f) != 0:
            e = self.err()
            self.L.flbgpu_filter_destroy(f)
            raise FlbGpuError("filter_init(%s): %s" % (plugin, e))
        return Filter(self, f)

    def chain(self, filters):
        c = self.L.flbgpu_chain_new(self.h)
        for f in filters:
            if self.L.flbgpu_chain_add(c, f.h) != 0:
                raise FlbGpuError("chain_add failed")
        if self.L.flbgpu_chain_init(c) != 0:
            raise FlbGpuError("chain_init: %s" % self.err())
        return Chain(self, c, filters)


class Parser:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def do(self, line):
        """flb_parser_do(): returns (ret, msgpack_bytes or None, (sec, nsec))."""
        out, n, t = C.c_void_p(), C.c_size_t(), Time()
        r = self.ctx.L.flbgpu_parser_do(self.h, line, len(line), C.byref(out), C.byref(n), C.byref(t))
        data = None
        if r >= 0 and out.value:
            data = C.string_at(out.value, n.value)
        if out.value:
            _libc.free(out)
        return r, data, (t.tv_sec, t.tv_nsec)


def _call_filter(fn, L, handle, data, tag):
    out, n = C.c_void_p(), C.c_size_t()
    buf = C.create_string_buffer(data, len(data)) if not isinstance(data, C.Array) else data
    r = fn(handle, C.cast(buf, C.c_void_p), len(data), _b(tag), len(_b(tag)), C.byref(out), C.byref(n))
    if r < 0:
        e = L.flbgpu_last_error()
        raise FlbGpuError("filter call failed: %s" % (e.decode(errors="replace") if e else "?"))
    res = None
    if r == FILTER_MODIFIED:
        res = C.string_at(out.value, n.value) if n.value else b""
    if out.value:
        _libc.free(out)
    return r, res


class Filter:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def cb(self, data, tag="test"):
        """cb_filter(): (FILTER_MODIFIED, bytes) or (FILTER_NOTOUCH, None)."""
        return _call_filter(self.ctx.L.flbgpu_filter_cb, self.ctx.L, self.h, data, tag)


class Chain:
    def __init__(self, ctx, h, filters):
        self.ctx, self.h, self.filters = ctx, h, filters

    def do(self, data, tag="test"):
        return _call_filter(self.ctx.L.flbgpu_chain_do, self.ctx.L, self.h, data, tag)

    def stats(self):
        s = Stats()
        self.ctx.L.flbgpu_chain_stats(self.h, C.byref(s))
        return s
