"""bench.py contract checks that need no GPU: the reference arm's JSON line (the driver parses it) and the fail-loud
behaviour of the product arm without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(args, timeout=240):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "kpoints_per_s_eigenval_fp64"
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["vs_baseline"] is None and line["higher_is_better"] is True and line["dtype"] == "f64"
    from oracle import ref_shim

    want_kind = "reference" if ref_shim.reference_available() else "port"
    assert line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["cores"] >= 1
    host = line["cpu_baseline"]["host"]
    assert host["numpy"] and host["scipy"] and "threadpool_info" in host
    for pool in host["worker_threadpool_info"] or []:  # BLAS threads = 1 in every worker (set before numpy is imported there)
        assert pool["num_threads"] == 1, pool
    assert line["scaling"] == "strong"
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"] and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1", "--nk", "1000"], timeout=120)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_reference_callable_is_the_same_model_as_the_packed_arrays():
    """bench.py builds a tbmodels.Model from the packed arrays for its CPU arms: same eigenvalues as the oracle."""
    import numpy as np

    import bench
    from oracle import ref_shim, tb_oracle as orc
    from oracle import workloads as wl

    if not ref_shim.reference_available():
        pytest.skip("reference package not available")
    tb = ref_shim.import_reference()
    for packed in (wl.haldane(), wl.synthetic(7, 9, seed=3)):
        fn, kind = bench._make_ref_callable(packed.R, packed.hop, packed.pos)
        assert kind == "reference"
        model = fn.__closure__[0].cell_contents if fn.__closure__ else None
        assert isinstance(model, tb.Model)
        k = np.random.default_rng(5).random((9, packed.dim))
        got = np.array(model.eigenval(k))
        want = orc.eigenval_array(packed.R, packed.hop, packed.pos, k)
        assert np.array_equal(got, want)
