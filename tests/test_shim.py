"""The drop-in boundary, end to end: the reference's OWN flb_filter_do() (src/flb_filter.c:119-323) drives the five
`struct flb_filter_plugin filter_gpu_*_plugin` of shim/filter_gpu.c -- compiled against the reference's headers into
oracle/_ref/flb-filter_gpu.so and registered the way src/flb_plugin.c registers a dynamic plugin -- and the result is
compared with the same configuration on the stock plugins: the chunk, byte for byte, the per-filter framework
counters flb_filter_do() keeps (filter_records_total, filter_bytes_total, filter_drop_records_total,
filter_drop_bytes_total, filter_add_records_total: src/flb_filter.c:221-303), Match routing, and the
log_to_metrics table.  CPU half: the plugins are bound to the CPU emulation of the device code; GPU half: to libflbgpu.so."""
import ctypes as C
import os
import re

import pytest

import cases
import l2m_cases
import util

pkg = util.pkg
TS = re.compile(r"^\S+Z ", re.M)
GPU_LIB = os.path.join(util.ROOT, "fluent-bit_b200", "libflbgpu.so")


def need_shim():
    if not (util.have_ref() and os.path.exists(util.SHIM_SO)):
        pytest.skip("oracle/_ref/flb-filter_gpu.so is not built")


def pipelines(lib_path, parsers, filters, match=None):
    """(stock reference pipeline, the same with every filter replaced by its gpu_ twin), each filter aliased f<k>"""
    out = []
    for gpu in (False, True):
        ref = util.Ref()
        if gpu:
            ref.load_gpu_plugins(lib_path)
        for kw in parsers:
            ref.parser(**kw)
        ins = []
        for k, (plugin, props) in enumerate(filters):
            extra = [("alias", "f%d" % k)]
            if match and k in match:
                extra.append(("match", match[k]))
            ins.append(ref.filter(("gpu_" if gpu else "") + plugin, list(props) + extra))
        out.append((ref, ins))
    return out


def same_behaviour(lib_path, parsers, filters, chunks, tag="test", match=None, l2m=None):
    (ref, rins), (gpu, gins) = pipelines(lib_path, parsers, filters, match)
    for chunk in chunks:
        assert gpu.chain_do(chunk, tag) == ref.chain_do(chunk, tag)
        for a, b in zip(rins, gins):
            assert gpu.filter_counters(b) == ref.filter_counters(a)
    if l2m is not None:
        shim = C.CDLL(util.SHIM_SO)
        shim.filter_gpu_l2m_text.restype = C.c_void_p
        shim.filter_gpu_l2m_text.argtypes = [C.c_void_p]
        p = shim.filter_gpu_l2m_text(gins[l2m])
        assert p
        got = C.string_at(p).decode(errors="replace")
        assert got == TS.sub("", ref.l2m_text(rins[l2m]))


CHAINS = ["north_star_chain", "json_chain_config1", "parser_modify_recmod", "grep_regex"]


def _chains(lib_path):
    for name in CHAINS:
        _, parsers, filters, mk = [c for c in cases.CASES if c[0] == name][0]
        chunk = mk()
        same_behaviour(lib_path, parsers, filters, [chunk, chunk[:len(chunk) // 2], chunk])
    # everything dropped by the first filter: the chain stops there and the drop counters take the whole chunk
    same_behaviour(lib_path, [cases.AP], [("grep", [("Regex", "log NOSUCHTEXT")]), ("modify", [("Add", "a b")])], [cases.apache_chunk(300)])
    # nothing touched at all
    same_behaviour(lib_path, [], [("grep", [("Exclude", "log NOSUCHTEXT")]), ("modify", [("Remove", "nosuchkey")])], [cases.apache_chunk(300)])


def _docker_decoders(lib_path):
    """the docker parser with its Decode_Field_As rules: the shim mirrors struct flb_parser, decoders included"""
    import test_decoders as T
    lines = [(b'{"log":"{\\"a\\":%d,\\"b\\":\\"x\\"}\\n","stream":"stdout","time":"2023-05-06T07:08:09.%dZ"}' % (i, i)) for i in range(200)]
    lines += [b'{"log":"plain %d\\n","stream":"stderr","time":"2023-05-06T07:08:10.5Z"}' % i for i in range(50)]
    same_behaviour(lib_path, [T.DOCKER], [("parser", [("Key_Name", "log"), ("Parser", "docker"), ("Reserve_Data", "On")]),
                                          ("grep", [("Regex", "stream stdout")])], [util.chunk_from_lines(lines)])


def _match_routing(lib_path):
    """Match decides per filter whether it sees the chunk (flb_router_match, src/flb_router.c)"""
    filters = [cases.P, ("grep", [("Regex", "method ^(GET|POST)$")]), ("modify", [("Add", "env prod")])]
    chunk = cases.apache_chunk(400)
    for tag in ("app.web", "app.db", "sys"):
        same_behaviour(lib_path, [cases.AP], filters, [chunk], tag=tag, match={0: "app.*", 1: "*.web", 2: "*"})


def _l2m(lib_path):
    for name in ("histogram_default_buckets", "counter_labels", "gauge_labels"):
        hit = [c for c in l2m_cases.L2M_CASES if c[0] == name]
        if not hit:
            continue
        _, parsers, filters, mk, k = hit[0]
        same_behaviour(lib_path, parsers, filters, [mk(), mk()], l2m=k)


def test_shim_exports(ref_available):
    need_shim()
    L = C.CDLL(util.SHIM_SO)
    for name in ("parser", "grep", "modify", "record_modifier", "log_to_metrics", "rewrite_tag", "multiline"):
        assert getattr(L, "filter_gpu_%s_plugin" % name)


def _rewrite_tag(lib_path):
    """BASELINE configs[3]: nginx parser + record_modifier + rewrite_tag -- the chunk that stays, the framework counters, and
    what reaches the emitter (in_emitter_add_record: per record there, per new tag here; the same bytes per tag)"""
    def per_tag(ref):
        d = {}
        for t, b in ref.emitted():
            d[t] = d.get(t, b"") + b
        return list(d.items())
    filters = [("parser", [("Key_Name", "log"), ("Parser", "nginx"), ("Reserve_Data", "On")]),
               ("record_modifier", [("Record", "hostname node-1"), ("Remove_key", "agent")]),
               ("rewrite_tag", [("Rule", "$code ^5 errors.$TAG[1].$method false"), ("Rule", "$method ^(PUT|HEAD)$ audit.$1.$TAG true"), ("Emitter_Name", "re_emitted")])]
    (ref, rins), (gpu, gins) = pipelines(lib_path, [cases.NG], filters)
    for seed in (1, 2):
        chunk = util.chunk_from_lines(util.apache_lines(400, seed=seed, nginx=True) + [b"not nginx"] * 2)
        ref.emit_reset()
        want = ref.chain_do(chunk, "web.front.access")
        want_emit = per_tag(ref)
        gpu.emit_reset()
        got = gpu.chain_do(chunk, "web.front.access")
        assert got == want
        assert per_tag(gpu) == want_emit and want_emit
        for a, b in zip(rins, gins):
            assert gpu.filter_counters(b) == ref.filter_counters(a)


def test_shim_rewrite_tag_hostsim(ref_available):
    need_shim()
    _rewrite_tag(util.HOSTSIM_SO)


@pytest.mark.gpu
def test_shim_rewrite_tag_gpu(gpu_lib, ref_available):
    need_shim()
    _rewrite_tag(GPU_LIB)


def test_shim_chains_hostsim(ref_available):
    need_shim()
    _chains(util.HOSTSIM_SO)


def test_shim_docker_decoders_hostsim(ref_available):
    need_shim()
    _docker_decoders(util.HOSTSIM_SO)


@pytest.mark.gpu
def test_shim_docker_decoders_gpu(gpu_lib, ref_available):
    need_shim()
    _docker_decoders(GPU_LIB)


def test_shim_match_routing_hostsim(ref_available):
    need_shim()
    _match_routing(util.HOSTSIM_SO)


def test_shim_l2m_hostsim(ref_available):
    need_shim()
    _l2m(util.HOSTSIM_SO)


def test_shim_refuses_what_the_stock_plugin_refuses(ref_available):
    need_shim()
    ref = util.Ref()
    ref.load_gpu_plugins(util.HOSTSIM_SO)
    with pytest.raises(RuntimeError):
        ref.filter("gpu_grep", [("Regex", "onlyonefield")])
    with pytest.raises(RuntimeError):
        ref.filter("gpu_parser", [("Key_Name", "log"), ("Parser", "does_not_exist")])


@pytest.mark.gpu
def test_shim_chains_gpu(gpu_lib, ref_available):
    need_shim()
    _chains(GPU_LIB)


@pytest.mark.gpu
def test_shim_match_routing_gpu(gpu_lib, ref_available):
    need_shim()
    _match_routing(GPU_LIB)


@pytest.mark.gpu
def test_shim_l2m_gpu(gpu_lib, ref_available):
    need_shim()
    _l2m(GPU_LIB)
