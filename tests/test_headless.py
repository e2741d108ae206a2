"""The headless C++ driver (fish_abm_b200/host): the reference's main() loop over the C ABI, with its two loggers."""
import os
import subprocess

import numpy as np
import pytest

from fish_abm_b200 import Ensemble, School

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "fish_abm_b200", "host", "fish_abm_headless")
QCSV = os.path.join(ROOT, "tests", "golden", "quality.csv")


def run(*args):
    return subprocess.run([EXE, *map(str, args)], capture_output=True, text=True, timeout=600)


def test_driver_is_built_and_links_only_the_c_abi():
    assert os.path.exists(EXE), "built by __graft_entry__.build()"
    r = run("--help")
    assert r.returncode == 0 and "--log-quality" in r.stdout
    src = open(os.path.join(ROOT, "fish_abm_b200", "host", "fishabm_host.cpp")).read() + open(os.path.join(ROOT, "fish_abm_b200", "host", "fish_abm_headless.cpp")).read()
    assert "cuda_runtime" not in src and "oracle" not in src.replace("// ", "")


def test_driver_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run("--quiet", "--quality", QCSV)
    assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--unfused"]])
def test_driver_logs_equal_the_python_binding(tmp_path, q100, extra):
    lq, li = tmp_path / "q.csv", tmp_path / "i.csv"
    r = run("--n", 60, "--steps", 130, "--seed", 11, "--quality", QCSV, "--log-quality", lq, "--log-intake", li, *extra)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "conserved: yes" in r.stdout
    dq = np.loadtxt(lq, delimiter=",", dtype=np.int64).reshape(1, -1)
    di = np.loadtxt(li, delimiter=",", dtype=np.int64)
    assert dq.shape == (1, 131) and dq[0, 0] == 3499 and di.shape == (13, 60)
    with School(n=60, seed=11, speed_norm=60.0) as S:
        S.initialize(5.0)
        S.set_quality(q100)
        sumq, intake = S.run_logged(130, intake_every=10)
    assert np.array_equal(dq[0, 1:], sumq)
    assert np.array_equal(di, intake)


@pytest.mark.gpu
def test_driver_ensemble_rows_have_the_shape_of_the_reference_data(tmp_path, q100):
    lq = tmp_path / "q.csv"
    r = run("--n", 60, "--steps", 200, "--seed", 5, "--replicates", 30, "--quality", QCSV, "--log-quality", lq, "--quiet")
    assert r.returncode == 0, r.stdout + r.stderr
    dq = np.loadtxt(lq, delimiter=",", dtype=np.int64)
    assert dq.shape == (30, 201)  # data/rolling_speed/*.csv: 30 replicates x (steps + 1) samples
    assert np.all(dq[:, 0] == 3499) and np.all(np.diff(dq, axis=1) <= 0)
    with Ensemble(30, n=60, seed=5, speed_norm=60.0) as E:
        E.initialize(5.0)
        E.set_quality(q100)
        sumq, _ = E.run_logged(200)
    assert np.array_equal(dq[:, 1:], sumq.T)


@pytest.mark.gpu
def test_driver_split_school_equals_the_unsplit_run(tmp_path, q100):
    """--slabs K: K slab contexts in the driver's process attached through peer memory; every log must be identical"""
    outs = []
    for k, extra in enumerate([[], ["--slabs", 2], ["--slabs", 4, "--devices", "0,0,0,0"]]):
        lq, li = tmp_path / f"q{k}.csv", tmp_path / f"i{k}.csv"
        r = run("--n", 3000, "--w", 2000, "--h", 2000, "--steps", 60, "--seed", 21, "--speed", 10, "--quality", QCSV, "--log-quality", lq,
                "--log-intake", li, "--quiet", *extra)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append((np.loadtxt(lq, delimiter=",", dtype=np.int64), np.loadtxt(li, delimiter=",", dtype=np.int64)))
    for q, i in outs[1:]:
        assert np.array_equal(q, outs[0][0]) and np.array_equal(i, outs[0][1])
    assert outs[0][0][0] == 3499 and outs[0][0][-1] < 3499
