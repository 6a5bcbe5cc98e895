"""Which property sets flb_filter_init() accepts: the same ones as the reference (oracle/_ref), and the accepted
ones give the same chunk.  Known, deliberate differences are listed: Uuid_key (random by definition) and three
modify conditions the reference accepts although it cannot evaluate them (missing argument, invalid regex)."""
import pytest

import cases
import util
from test_chain_parity import run_case

pkg = util.pkg
CONFIGS = {
    "grep": [[("Regex", "log")], [("Regex", "log  GET")], [("Regex", " log GET")], [("Regex", "log GET extra")], [("Regex", 'log "GET /"')], [("Regex", "")],
             [("Exclude", "log")], [("Logical_Op", "xor"), ("Regex", "log G")], [("regex", "LOG get")], [("Regex", "log [")], [("Regex", "log (")], [("Regex", "$log GET")],
             [("Regex", "$log['a' GET")], [("Regex", "log /GET/i")], [("Regex", "log /get/i")], [("Regex", "log /G E T/x")], [("Logical_Op", "and"), ("Logical_Op", "or"), ("Regex", "log G")]],
    "parser": [[("Key_Name", "log")], [("Parser", "apache")], [("Key_Name", "log"), ("Parser", "nope")], [("Key_Name", "log"), ("Parser", "apache"), ("Reserve_Data", "maybe")],
               [("Key_Name", ""), ("Parser", "apache")], [("Key_Name", "log"), ("Parser", "apache"), ("Unescape_Key", "on")], [("Key_Name", "log"), ("Parser", "apache"), ("Preserve_Key", "yes")],
               [("Key_Name", "log"), ("Parser", "apache"), ("Preserve_Key", "1")], [("Key_Name", "log"), ("Parser", "apache"), ("Reserve_Data", "TRUE")],
               [("key_name", "log"), ("parser", "apache"), ("reserve_data", "off")], [("Key_Name", "log"), ("Key_Name", "level"), ("Parser", "apache")],
               [("Key_Name", "log"), ("Parser", "apache"), ("Reserve_Data", "on"), ("Reserve_Data", "on")]],
    "record_modifier": [[("Record", "a")], [("Record", "a b c")], [("Record", "a  b")], [("Record", 'a "b c"')], [("Remove_key", "")], [("Allowlist_key", "*")], [("Remove_key", "l*g")],
                        [("Record", "")], [("Whitelist_key", "log"), ("Allowlist_key", "level")], [("record", "a b"), ("remove_key", "LOG")], [("Allowlist_key", ""), ("Allowlist_key", "log")]],
    "modify": [[("Set", "a")], [("Set", "a b c d")], [("Remove", "a b")], [("Rename", "a")], [("Condition", "Nope a")], [("Condition", "Key_value_equals a")], [("Set", 'a "b c"')],
               [("Set", '"a b" c')], [("Add", "level x"), ("add_if_not_present", "lvl y")], [("Condition", "key_exists level"), ("condition", "KEY_EXISTS log"), ("SET", "z 1")],
               [("Remove_regex", "[")], [("Remove_regex", "")], [("Condition", "Matching_keys_have_matching_values l"), ("Add", "a b")], [("Move_to_start", "")], [("Copy", "a b c")]],
    "log_to_metrics": [[("metric_description", "d"), ("tag", "t"), ("metric_mode", "counter"), ("metric_mode", "gauge"), ("value_field", "n")],
                       [("metric_description", "d"), ("tag", "t"), ("tag", "u")], [("metric_description", "d"), ("tag", "t"), ("bucket", "1"), ("bucket", "1")],
                       [("metric_description", "d"), ("tag", "t"), ("kubernetes_mode", "maybe")], [("metric_description", "d"), ("tag", "t"), ("discard_logs", "maybe")],
                       [("metric_description", "d"), ("tag", "t"), ("add_label", "a b c")], [("metric_description", "d"), ("tag", "t"), ("metric_mode", "COUNTER")],
                       [("metric_description", ""), ("tag", "t")], [("tag", "t")]],
}
ALL = [(plugin, props) for plugin, sets in CONFIGS.items() for props in sets]


@pytest.mark.parametrize("plugin,props", ALL, ids=["%s-%d" % (p, i) for p, sets in CONFIGS.items() for i in range(len(sets))])
def test_same_configurations_are_accepted(plugin, props, sim_lib, ref_available):
    ref, ctx = util.Ref(), pkg.Context(0, lib=sim_lib)
    ref.parser(**cases.AP)
    ctx.parser(**cases.AP)
    try:
        ref.filter(plugin, props)
        accepted = True
    except RuntimeError:
        accepted = False
    if not accepted:
        with pytest.raises(pkg.FlbGpuError):
            ctx.filter(plugin, props)
        return
    ctx.filter(plugin, props)                                 # accepted here too
    if plugin != "log_to_metrics":
        recs = util.split_records(cases.mixed_chunk())
        run_case(sim_lib, [cases.AP], [(plugin, props)], cases.mixed_chunk()[:recs[30][0]])


PARSERS = [dict(name="a", format="regex"), dict(name="a", format="regex", regex=""), dict(name="a", format="regex", regex="("), dict(name="a", format="nope"), dict(name="a", format="JSON"),
           dict(name="a", format="json", time_fmt="%Y"), dict(name="a", format="json", time_key="t"), dict(name="a", format="json", time_fmt="%Q", time_key="t"),
           dict(name="a", format="json", time_fmt="%Y-%m-%d %H:%M:%S.%L %z", time_key="t", time_offset="+0100"), dict(name="a", format="json", time_fmt="%H", time_key="t", time_offset="bad"),
           dict(name="a", format="json", time_fmt="%H", time_key="t", time_offset="+01:00"), dict(name="a", format="json", time_fmt="%H", time_key="t", time_offset="0100"),
           dict(name="a", format="regex", regex="^(?<x>.)$", types="x:integer y:float"), dict(name="a", format="regex", regex="^(?<x>.)$", types="x:nope"),
           dict(name="", format="json"), dict(name="a", format="ltsv", regex="ignored"), dict(name="a", format="logfmt", time_fmt="%s", time_key="t", time_keep=True),
           dict(name="a", format="regex", regex="^(.)$"), dict(name="a", format="regex", regex="^(?<x>.)(?<x>.)$"), dict(name="a", format="regex", regex="/^(?<x>.)$/i"),
           dict(name="a", format="regex", regex="^(?<time>.*)$", time_fmt="%L", time_key="time"), dict(name="a", format="regex", regex="^(?<time>.*)$", time_fmt="%Y %L %L", time_key="time"),
           dict(name="a", format="regex", regex="^(?<time>.*)$", time_fmt="%b %d %H:%M:%S", time_key="time"), dict(name="a", format="regex", regex="^(?<time>.*)$", time_fmt="%Y %q", time_key="time")]
LINES = [b"x", b"xy", b'{"t":"2023","a":1}', b"t=5 a=1", b"t:5\ta:1", b"2023 123 456", b"Feb  3 04:05:06", b"123", b"2023 q"]


@pytest.mark.parametrize("kw", PARSERS, ids=[str(i) for i in range(len(PARSERS))])
def test_same_parser_definitions_are_accepted(kw, sim_lib, ref_available):
    """flb_parser_create(): same definitions accepted; a conversion neither strptime knows is accepted and fails per line"""
    ref, ctx = util.Ref(), pkg.Context(0, lib=sim_lib)
    try:
        rp = ref.parser(**kw)
    except RuntimeError:
        with pytest.raises(pkg.FlbGpuError):
            ctx.parser(**kw)
        return
    p = ctx.parser(**kw)
    for line in LINES:
        r, data, t = p.do(line)
        rr, rdata, rt = ref.parser_do(rp, line)
        assert (r < 0) == (rr < 0), line
        if rr >= 0:
            assert data == rdata and t == rt, line
