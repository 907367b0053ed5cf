"""Host-side pieces of bench.py that need no GPU: the algorithmic byte count of the fused kernel, the committed DRAM
traffic and access-pattern ceiling it reports beside the roofline, and the reference arm's argument handling."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_algorithmic_bytes_of_config2():
    import bench
    # 32 (pose) + 8 (slot, aux) + 4n (keys) + 68e (records read) + 68m (written) + 8 (weight) + 4K (ids)
    b = bench.k2_bytes_per_particle("f32", 64, 8, 8.0, 1.0)
    assert b == 32 + 8 + 4 * 64 + 68 * 8 + 68 * 8 + 8 + 4 * 8 == 1424
    b64 = bench.k2_bytes_per_particle("f64", 64, 8, 8.0, 1.0)
    assert b64 == 32 + 8 + 256 + 164 * 8 + 164 * 8 + 8 + 32


def test_committed_traffic_matches_the_benchmarked_workload():
    import bench
    for arith, dtype, key in (("f32", "f32", "measure_kernel<float,f32>"), ("f64", "f32", "measure_kernel<float>"),
                              ("f64", "f64", "measure_kernel<double>")):
        traffic, src = bench.traffic_from_profiles(arith, dtype, 1 << 20, 64, 8)
        assert src == "r2_traffic.json" and traffic > 0, key
        algorithmic = bench.k2_bytes_per_particle(dtype, 64, 8, 8.0, 1.0) * (1 << 20)
        # the fp32-algebra kernel (the bench default) moves at most 1.15x its algorithmic bytes; the fp64 ones 1.25x
        assert traffic / algorithmic < (1.15 if arith == "f32" else 1.25), (key, traffic / algorithmic)
    # another workload: no committed capture applies
    assert bench.traffic_from_profiles("f32", "f32", 1 << 19, 64, 8) == (None, None)


def test_pattern_ceiling_is_reported_for_the_probed_workload_only():
    import bench
    args = argparse.Namespace(dtype="f32", arith="f32")
    pc = bench.pattern_ceiling(args, 1 << 20, 64, 8, 0.43)
    assert pc is not None and 0.3 < pc["ms"] < 0.5 and abs(pc["k2_over_ceiling"] - 0.43 / pc["ms"]) < 1e-12
    assert bench.pattern_ceiling(args, 1 << 20, 256, 8, 2.0) is None
    assert bench.pattern_ceiling(argparse.Namespace(dtype="f64", arith="f64"), 1 << 20, 64, 8, 1.0) is None
    with open(os.path.join(ROOT, "profiles", "r2_k2_mem_probe.json")) as fh:
        pr = json.load(fh)
    # the probe's own numbers: reads alone are cheap, the write-backs are what costs
    assert pr["reads_only_ms"] < 0.5 * pr["all_ms"] < pr["records_read_write_ms"]


def test_bench_lines_in_profiles_carry_the_contract_keys():
    for name in ("r2_bench_n1.json", "r2_bench_n2.json", "r2_bench_n8.json"):
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            line = json.loads(fh.read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
            assert key in line, (name, key)
        assert line["metric"] == "particle_observation_updates_per_sec" and line["scaling"] == "weak"
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
        assert "workload" in line["config"]
        if line["n_gpus"] > 1:
            assert line["sharded_identical"] is True
        else:
            assert line["cpu_baseline"]["kind"] == "reference"
            assert line["e2e"]["ms_per_step"] >= 0.999 * line["ms_per_step"]   # e2e is a superset of the device loop
