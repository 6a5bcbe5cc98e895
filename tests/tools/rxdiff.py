"""Differential check of the device regex VM (CPU emulation, tests/hostsim) against the
real Onigmo through oracle/_ref/libflbref.so.  Dev tool + imported by tests."""
import ctypes as C, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = C.CDLL(os.path.join(ROOT, 'oracle/_ref/libflbref.so'))
SIM = C.CDLL(os.path.join(ROOT, 'tests/hostsim/libhostsim.so'))
REF.flbref_regex_create.restype = C.c_void_p
REF.flbref_regex_create.argtypes = [C.c_char_p]
REF.flbref_regex_search.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
REF.flbref_regex_match.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
REF.flbref_regex_names.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
REF.flbref_regex_destroy.argtypes = [C.c_void_p]
SIM.sim_rx_compile.restype = C.c_void_p
SIM.sim_rx_compile.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
SIM.sim_rx_search.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_uint]
SIM.sim_rx_free.argtypes = [C.c_void_p]
SIM.sim_rx_ngroups.argtypes = [C.c_void_p]
SIM.sim_rx_names.argtypes = [C.c_void_p, C.c_char_p, C.c_int]

def ref_search(re, s):
    beg = (C.c_int * 64)(); end = (C.c_int * 64)(); n = C.c_int()
    m = REF.flbref_regex_match(re, s, len(s))
    r = REF.flbref_regex_search(re, s, len(s), beg, end, 64, C.byref(n))
    if m <= 0:
        return None
    if r <= 0:   # matched but zero groups: region was freed; only know it matched
        return ('m',)
    return tuple((beg[i], end[i]) for i in range(n.value))

def sim_search(h, s, stack=4096, budget=2000000):
    ng = SIM.sim_rx_ngroups(h)
    caps = (C.c_int * (2 * (ng + 1)))()
    r = SIM.sim_rx_search(h, s, len(s), caps, stack, budget)
    if r == 0:
        return None
    if r < 0:
        return ('err', r)
    return tuple((caps[2 * i], caps[2 * i + 1]) for i in range(ng + 1))

def compare(pattern, subjects, verbose=True):
    """returns (n_checked, n_bad, status) status in ok/ref_reject/sim_reject"""
    pb = pattern if isinstance(pattern, bytes) else pattern.encode()
    re = REF.flbref_regex_create(pb)
    err = C.create_string_buffer(200)
    h = SIM.sim_rx_compile(pb, err, 200)
    if not re and not h:
        return 0, 0, 'both_reject'
    if not re:
        if verbose: print('SIM accepts but REF rejects:', pb)
        SIM.sim_rx_free(h)
        return 0, 1, 'ref_reject'
    if not h:
        REF.flbref_regex_destroy(re)
        return 0, 0, 'sim_reject:' + err.value.decode()
    bad = 0
    for s in subjects:
        a = ref_search(re, s); b = sim_search(h, s)
        if a == ('m',):
            ok = b is not None and b[0] != 'err'
        else:
            ok = (a == b)
        if not ok:
            bad += 1
            if verbose: print('MISMATCH pat=%r subj=%r\n  ref=%r\n  sim=%r' % (pb, s, a, b))
    REF.flbref_regex_destroy(re); SIM.sim_rx_free(h)
    return len(subjects), bad, 'ok'

ATOMS = ['a', 'b', 'c', ' ', '.', r'\d', r'\w', r'\s', r'\S', '[a-c]', '[^ ]', '[^"]', 'x', '-', '"', r'\[', r'\]', '[ab]', r'\W', 'é', '[^a]', r'\.', ':', '0']
def rand_pat(rng, depth=0):
    n = rng.randint(1, 4)
    parts = []
    for _ in range(n):
        r = rng.random()
        if r < 0.55 or depth > 2:
            a = rng.choice(ATOMS)
        elif r < 0.70:
            a = '(?:' + rand_alt(rng, depth + 1) + ')'
        elif r < 0.90:
            a = '(?<g%d>' % rng.randint(0, 99999) + rand_alt(rng, depth + 1) + ')'
        elif r < 0.95:
            a = rng.choice(['^', '$', r'\b', r'\A', r'\z', r'\Z', r'\B'])
        else:
            a = rng.choice(['(?=' , '(?!', '(?>']) + rand_alt(rng, depth + 1) + ')'
        q = rng.random()
        if a not in ('^', '$', r'\b', r'\A', r'\z', r'\Z', r'\B') and not a.startswith('(?=') and not a.startswith('(?!'):
            if q < 0.12: a += '*'
            elif q < 0.20: a += '+'
            elif q < 0.28: a += '?'
            elif q < 0.33: a += '*?'
            elif q < 0.37: a += '+?'
            elif q < 0.40: a += '??'
            elif q < 0.44: a += '{%d,%d}' % (rng.randint(0, 2), rng.randint(2, 4))
            elif q < 0.46: a += '{%d}' % rng.randint(1, 3)
            elif q < 0.48: a += '{%d,}' % rng.randint(0, 2)
            elif q < 0.50: a += rng.choice(['*+', '++', '?+'])
            elif q < 0.52: a += '{1,3}?'
        parts.append(a)
    return ''.join(parts)

def rand_alt(rng, depth):
    k = 1 if rng.random() < 0.7 else rng.randint(2, 3)
    return '|'.join(rand_pat(rng, depth) for _ in range(k))

ALPH = ['a', 'b', 'c', ' ', 'x', '1', '0', '-', '"', '[', ']', '.', ':', '\n', 'é', '_', 'a', 'b', ' ', '\t']
def rand_subj(rng):
    return ''.join(rng.choice(ALPH) for _ in range(rng.randint(0, 14))).encode()

def fuzz(seed, npat, nsub=40, verbose=True):
    rng = random.Random(seed)
    tot = bad = rej = 0
    rejects = {}
    for _ in range(npat):
        p = rand_alt(rng, 0)
        subs = [rand_subj(rng) for _ in range(nsub)]
        n, b, st = compare(p, subs, verbose)
        tot += n; bad += b
        if st.startswith('sim_reject'):
            rej += 1; rejects[st] = rejects.get(st, 0) + 1
    return tot, bad, rej, rejects

if __name__ == '__main__':
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    npat = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    print(fuzz(seed, npat))
