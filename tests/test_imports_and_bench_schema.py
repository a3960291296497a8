"""CPU: every module of the package imports without a GPU (the CUDA library only loads; no kernel runs), and the
reference arm of bench.py prints the contract's keys."""
import importlib
import io
import json
import os
import sys
from contextlib import redirect_stdout

import instantavatar_b200


def test_every_module_imports_on_cpu():
    root = os.path.dirname(instantavatar_b200.__file__)
    names = []
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py") and f != "__init__.py":
                rel = os.path.relpath(os.path.join(dirpath, f), os.path.dirname(root))
                names.append(rel[:-3].replace(os.sep, "."))
    assert len(names) >= 15, names
    for name in sorted(names):
        importlib.import_module(name)


def test_reference_arm_json_contract(monkeypatch):
    sys.path.insert(0, ".")
    import bench

    class FakeFrame:  # the real one renders a 512x512 frame on the CPU oracle (seconds); the schema does not need it
        threads = 3
        def __init__(self, frame):
            pass
        def step(self):
            pass

    monkeypatch.setattr(bench, "CpuFrame", FakeFrame)
    args = bench.argparse.Namespace(gpus=1, steps=2, warmup=1)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["config"]["workload"]
    assert line["cpu_baseline"]["cores"] == 3 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # other ranks of a torchrun launch print nothing
    monkeypatch.setenv("RANK", "1")
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    assert buf.getvalue() == ""
