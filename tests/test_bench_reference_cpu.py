"""CPU: `bench.py --impl reference` (the reference arm the driver runs beside ours) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1",
                          "--warmup", "1", "--no-encode"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "hamming_retrieval_query_x_gallery_pairs_per_sec"
    assert line["unit"] == "pairs/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["config"]["workload"] == "C1" and line["steps"] == 1 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_flop_count_matches_the_oracle_and_the_product_arm_does_not_import_it():
    """bench.py's encoder FLOP count is its own (8.818 GFLOP per image) and equals the oracle's; outside the
    cpu_baseline / reference legs bench.py never touches oracle/ (the product arm must not route through the checker)."""
    import ast
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from oracle import clip_port

    assert bench.encode_flops_per_image() == clip_port.flops_image() == 8817623040
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_reference_map", "cpu_reference_topk", "cpu_encode_images_per_sec", "reference_cuda_encode"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle" for n in ast.walk(fn))
        uses |= any(isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names) for n in ast.walk(fn))
        assert not uses or fn.name in allowed, fn.name
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in ast.dump(n) for n in top)


def test_product_package_never_imports_the_oracle():
    import ast

    pkg = os.path.join(ROOT, "clip_based_cross_modal_hash_b200")
    for name in sorted(os.listdir(pkg)):
        if not name.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, name)).read())
        for n in ast.walk(tree):
            if isinstance(n, ast.ImportFrom):
                assert (n.module or "").split(".")[0] != "oracle", name
            if isinstance(n, ast.Import):
                assert all(a.name.split(".")[0] != "oracle" for a in n.names), name
