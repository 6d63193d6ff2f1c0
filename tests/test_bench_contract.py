"""CPU checks of the bench.py JSON contract on the committed bench lines (profiles/): the keys the driver reads, the
roofline / cpu_baseline / e2e objects, and that the reference arm prints the same shape."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        return json.load(fh)


def test_bench_line_has_every_contract_key():
    d = _line("r02_bench_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "images/sec" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    assert d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] == 5 * 4 * 512 * 1024 * 4 and d["e2e"]["value"] < d["value"] * 1.02
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # round 2: fixed-protocol CPU leg and the stock PyTorch / cuDNN legs travel in the same line
    assert "1 full 512x1024 image per step" in c["sample"] and "1 warm-up + 2 timed" in c["sample"]
    t = d["torch_gpu"]
    for variant in ("tf32", "bf16_channels_last"):
        assert t[variant]["cudnn_benchmark"] is True and t[variant]["warmup"] >= 5 and t[variant]["steps"] >= 10
        assert t[variant]["ms_per_step"] > 0
    assert abs(t["speedup_vs_tf32"] - t["tf32"]["ms_per_step"] / d["ms_per_step"]) < 1e-9
    assert d["cuda_graph"] is True and d["peak_mem_gb"] > 1


def test_side_lines_for_configs_4_and_5():
    for name, px, batch in (("r02_bench_config4_local_enhancer.json", "1024x2048", 1), ("r02_bench_config5_box2mask.json", "256x256", 8)):
        d = _line(name)
        assert px in d["metric"] and d["config"]["per_gpu_batch"] == batch and d["gpu_launches"] > 0
        assert abs(d["value"] - batch * d["n_gpus"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] <= d["value"] * 1.02


def test_two_gpu_lines_check_the_replicas():
    for name, n in (("r02_bench_2gpu_buckets8.json", 2), ("r02_bench_2gpu_single.json", 2), ("r02_bench_8gpu_buckets8.json", 8),
                    ("r02_bench_8gpu_single.json", 8)):
        d = _line(name)
        assert d["n_gpus"] == n and d["replicas_identical"] is True and d["cuda_graph"] is True
        assert d["config"]["global_batch"] == 4 * n and abs(d["value"] - 4 * n / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    with open(os.path.join(ROOT, "profiles", "r02_multigpu_check.json")) as fh:
        c = json.load(fh)
    assert c["replicas_identical"] is True and c["graph"] is True and c["loss_err"][0] < 1e-5 and max(c["loss_err"]) < 5e-3


def test_multi_gpu_lines_report_whole_job_throughput():
    for name, n in (("r01_bench_2gpu_v3.json", 2), ("r01_bench_8gpu_bf16x3_v3.json", 8), ("r01_bench_8gpu_bf16_v3.json", 8)):
        d = _line(name)
        assert d["n_gpus"] == n and d["config"]["global_batch"] == 4 * n and d["config"]["parallelism"] == "dp%d" % n
        assert abs(d["value"] - 4 * n / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]


def test_k1_traffic_file_feeds_the_roofline():
    with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as fh:
        t = json.load(fh)
    for prec in ("bf16x3", "bf16"):
        assert t[prec]["dram_bytes_read"] > 0 and t[prec]["dram_bytes_write"] > 0
        assert os.path.exists(os.path.join(ROOT, t[prec]["source"]))
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    tr, src = bench.k1_traffic(True)
    assert tr == t["bf16x3"]["dram_bytes_read"] + t["bf16x3"]["dram_bytes_write"] and src.endswith(".txt")
    # algorithmic bytes of K1 in bf16x3: hi+lo operands (activations with border + weights) + fp32 output
    algo = 2 * (4 * 34 * 66 * 1024 * 2) + 2 * (9 * 1024 * 1024 * 2) + 4 * 32 * 64 * 1024 * 4
    assert 0.8 * algo < tr < 1.3 * algo
