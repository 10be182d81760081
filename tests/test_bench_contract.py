"""The JSON line of bench.py (committed copies under profiles/) carries every key of the measurement contract;
bench.py itself parses and refuses to run the B200 arm without a GPU (no CPU fallback)."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_bench.json")))


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_lines_follow_the_contract(path):
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not (set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]


def test_last_line_has_native_and_pic_legs():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01n_bench.json")))
    assert d["native"]["field_step_ms"] > 0 and d["native"]["atoms_interp_per_s"] > 0 and d["native"]["cpu_baseline"]["cores"] >= 1
    assert d["pic"]["ms_per_step"] > 0 and d["pic"]["particles_alive"] > 0


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_synthetic_pic_workload_is_consistent(golden):
    """config 3 particles of bench.py: every electron lies in the solver cell it claims (checked with the oracle's
    reference-pinned LinearHexahedra search), within 50 A + a cell of the apex, deterministic for a seed"""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle.oracle import Oracle
    m = golden("mesh", "mdsmall")
    pos, vel, cells = bench.pic_particles(m, 3000)
    pos2, _, cells2 = bench.pic_particles(m, 3000)
    assert np.array_equal(pos, pos2) and np.array_equal(cells, cells2)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    assert np.array_equal(o.particle_cells(pos, cells), cells)
    apex = m["surf_atoms"][np.argmax(m["surf_atoms"][:, 2])]
    assert np.linalg.norm(pos - apex, axis=1).max() < 60.0
    assert abs(vel.std() - 0.1) < 0.01


def test_synthetic_x_particles_lie_in_their_cells():
    """config 4 electrons of bench.py (trilinear images of natural coordinates): inside the hexahedron they name"""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from femocs_b200 import synth
    nodes, hexs, mk = bench.load_x_mesh(0)
    pxyz, pcell = bench.synth_particles(nodes, hexs, 2000)
    lo = nodes[hexs[pcell]].min(1); hi = nodes[hexs[pcell]].max(1)
    assert np.all(pxyz >= lo - 1e-9) and np.all(pxyz <= hi + 1e-9) and pcell.min() >= 0 and pcell.max() < len(hexs)
