"""ctypes access to the regex VM in the CPU emulation library (no reference needed)."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SIM = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
SIM.sim_rx_compile.restype = C.c_void_p
SIM.sim_rx_compile.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
SIM.sim_rx_search.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_uint]
SIM.sim_rx_ngroups.argtypes = [C.c_void_p]


def compile(pattern):
    err = C.create_string_buffer(200)
    return SIM.sim_rx_compile(pattern.encode() if isinstance(pattern, str) else pattern, err, 200)


def search(h, s, stack=4096, budget=4000000):
    ng = SIM.sim_rx_ngroups(h)
    caps = (C.c_int * (2 * (ng + 1)))()
    r = SIM.sim_rx_search(h, s, len(s), caps, stack, budget)
    if r == 0:
        return None
    if r < 0:
        return ("err", r)
    return tuple((caps[2 * i], caps[2 * i + 1]) for i in range(ng + 1))
