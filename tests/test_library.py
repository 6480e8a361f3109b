"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol that
include/gpsurv.h declares; without a GPU it fails loudly instead of falling back; the product
package never imports the oracle."""
import ast
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gpsurv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gps_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from gpsurvival_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in gpsurv.h but not exported by libgpsurv.so"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names


def test_version_and_loud_failure_without_gpu(lib):
    assert b"sm_100a" in lib.gps_version()
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    rc = lib.gps_create(0, C.byref(h))
    assert rc == 2 and not h                       # GPS_ERR_CUDA, no handle, no CPU fallback
    assert b"no CPU fallback" in lib.gps_last_error(None)
    from gpsurvival_b200 import gp
    import pytest
    with pytest.raises(gp.GpsError):
        gp.ForOptimisationLog(np.zeros(3), np.zeros((4, 1)), np.zeros(4), np.ones(4), "Zero", 1, None, "SqExp", 4,
                              False, None, False)


def test_null_handle_is_an_argument_error(lib):
    assert lib.gps_nlml(None, None, 0, None) == 1
    assert lib.gps_destroy(None) == 0
    assert lib.gps_launch_count(None) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gpsurvival_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(dirpath, f)).read())
                for node in ast.walk(tree):
                    mods = []
                    if isinstance(node, ast.Import):
                        mods = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        mods = [node.module or ""]
                    assert not any(m.split(".")[0] == "oracle" for m in mods), f"{f} imports the oracle"
            if f.endswith((".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read()


def test_host_logic_flatten_and_options():
    from gpsurvival_b200 import gp, synth
    lh = synth.start_log_hyp(3, "ARD")
    assert gp._flatten(lh, 3, "Zero", False, None).size == 5
    assert gp._flatten(lh, 3, "Linear", False, None).size == 9
    assert gp._flatten(lh, 3, "Zero", "noiseCorrLearned", [0.1]).size == 6
    cfg = synth.CONFIGS["C2"]
    assert cfg.n_hyp == 22 and synth.start_par(cfg).size == 22
    pb = synth.make_config_problem(cfg, scale=0.05)
    assert pb["trainingData"].shape == (100, 20) and int(np.sum(pb["events"] == 0)) == 50
    assert abs(pb["trainingData"].mean()) < 1e-12 and np.allclose(pb["trainingData"].std(axis=0, ddof=1), 1.0)


def test_r_glue_type_checks_against_the_abi():
    """R is not installed here, so r/gpsurv_glue.c can never be run; it can still be TYPE-CHECKED:
    gcc -fsyntax-only against a stub of the few R C-API declarations it uses (tests/r_stub) and the real
    include/gpsurv.h — every gps_* call in the glue must match the ABI's argument list, and the .Call
    registration table must state each wrapper's true argument count."""
    import re
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    glue = os.path.join(root, "r", "gpsurv_glue.c")
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the build image"
    res = subprocess.run([gcc, "-fsyntax-only", "-std=c11", "-Wall", "-Werror=implicit-function-declaration",
                          "-Werror=incompatible-pointer-types", "-Werror=int-conversion", "-Wno-cast-function-type",
                          "-I", os.path.join(root, "tests", "r_stub"), glue], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    src = open(glue).read()
    defs = {m.group(1): m.group(2).count("SEXP") for m in re.finditer(r"^SEXP (gps_r_\w+)\(([^)]*)\)", src, re.M)}
    table = re.findall(r'\{"(gps_r_\w+)", \(DL_FUNC\)&(gps_r_\w+), (\d+)\}', src)
    assert table and len(table) == len(defs)
    for name, fn, nargs in table:
        assert name == fn and defs[fn] == int(nargs), (name, defs.get(fn), nargs)


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the reference algorithm (numpy restatement) on the host cores, one JSON line with
    the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "gp_logml_grad_evals_per_sec" and line["unit"] == "evals/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
