"""CPU checks of bench.py's bookkeeping (no GPU): the step sizing, the config both arms print, and the ncu-traffic stamp."""
import importlib.util
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_a_step_is_at_least_two_thousand_passes_in_total():
    for steps in (1, 3, 20, 100, 2000, 5000):
        r = bench.passes_per_step(steps)
        assert r >= 1 and r * steps >= 2000 and (r - 1) * steps < 2000 or r == 1       # >= ~2.5 s at 1.3 ms per pass
    assert bench.passes_per_step(20, override=7) == 7


def test_both_arms_print_the_same_config():
    a = bench.bench_config(8, bench.passes_per_step(20))
    b = bench.bench_config(8, bench.passes_per_step(20), P_mean=1.0)                    # the b200 arm adds measured extras
    assert a["workload"] == b["workload"] == bench.WORKLOAD
    assert all(a[k] == b[k] for k in ("views_per_pass_per_gpu", "passes_per_step", "views_per_step_per_gpu"))
    assert a["views_per_step_per_gpu"] == 8 * 100 and "model" not in a


def test_traffic_stamp_matches_code_not_comments():
    h = bench._source_hash("render_forward_kernel")
    assert re.fullmatch(r"[0-9a-f]{64}", h) and h != bench._source_hash("render_backward_q_kernel")
    with open(os.path.join(ROOT, "profiles", "render_forward_traffic.json")) as f:
        j = json.load(f)
    for kernel in ("render_forward_kernel", "render_backward_q_kernel"):
        val, src = bench._traffic(kernel)
        if j[kernel]["source_sha256"] == bench._source_hash(kernel):
            assert val == j[kernel]["dram_bytes_per_launch"] and 3e7 < val < 3e8         # tens of MB per launch at C2
        else:                                                                             # kernel code changed since the capture:
            assert val is None and src.startswith("stale")                                # refused, never silently reported
    assert bench._traffic("no_such_kernel")[0] is None
