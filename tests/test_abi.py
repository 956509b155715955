"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/plonky2_b200.h declares, and refuses to compute without a CUDA device (no CPU path)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            names |= set(re.findall(r"\beng_status\s+(eng_\w+)\s*\(", src))
    return sorted(names)


def test_header_compiles_as_c():
    # the boundary is plain C: no C++ or torch types in the signatures
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "plonky2_b200.h")])


def test_library_exports_every_declared_symbol():
    import eth_lc_plonky2_b200 as E
    E.build()
    lib = ctypes.CDLL(E.so_path())
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), "libplonky2_b200.so does not export %s" % name
    assert set(E.exported_symbols()) == set(decl)     # the Python binding covers the whole header


def test_product_does_not_touch_oracle():
    pkg = os.path.join(ROOT, "eth-lc-plonky2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in src.replace("never includes anything from oracle/", "") or f == "poseidon_consts.h" or "from oracle" not in src
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import eth_lc_plonky2_b200 as E
    with pytest.raises(E.EngineError) as ei:
        E.init()
    assert ei.value.status == E.ENG_ERR_STATE
    with pytest.raises(E.EngineError):
        E.poseidon([0] * 12)                     # no silent CPU fallback


def test_rust_sys_declarations_match_the_header():
    """integration/plonky2-b200-sys/src/lib.rs is generated from include/plonky2_b200.h (tools/gen_rust_sys.py): the
    reference-side binding cannot drift from the C ABI."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"])
    assert r.returncode == 0, "run python tools/gen_rust_sys.py"


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys,
    the sample it really ran and a zero-copy e2e object.  Tiny workload here; no GPU involved."""
    import json
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--log-n", "10"], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "commit_throughput" and line["unit"] == "GB/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["config"]["sample_log_rows"] <= 10 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
