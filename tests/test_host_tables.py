"""CPU: host-side table builders and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from gym_po import _native as N
    header = open(os.path.join(ROOT, "include", "gpt_b200.h")).read()
    declared = set(re.findall(r"^GPT_API [\w\s\*]+?\b(gpt_\w+)\(", header, flags=re.M))
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    lib = ctypes.CDLL(os.path.abspath(N.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert N.lib.gpt_abi_version() == N.ABI_VERSION


def test_config_struct_layout_matches_header():
    """ctypes mirror of gpt_config: same field names, same order as the C header."""
    from gym_po import _native as N
    header = open(os.path.join(ROOT, "include", "gpt_b200.h")).read()
    start = header.index("typedef struct gpt_config {") + len("typedef struct gpt_config {")
    body = header[start:header.index("} gpt_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        for part in stmt.split(","):
            names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    assert names == [f[0] for f in N.GptConfig._fields_]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_po.envs import TaxiVecEnv
    with pytest.raises(RuntimeError):
        TaxiVecEnv(8)


@pytest.mark.parametrize("m", [oracle.TAXI_MAP, oracle.EXTENDED_TAXI_MAP,
                               ("A: |B:C", " : : | ", "D| : :E")])
def test_taxi_wall_bits_equal_the_reference_motion_rule(m):
    """bit a of wall_bits[cell] == 'move a is blocked' under the oracle's character-map rule
    (target is '|' or a '|' separator is crossed, extended_taxi.py:248-260), for every cell x action;
    and wall_bits == the oracle's hansen_encodings; locations and valid states agree."""
    from gym_po.envs.extended_taxi import parse_taxi_map
    rows, cols, wall_bits, locs, is_wall = parse_taxi_map(m)
    orc = oracle.TaxiOracle(1, map=m)
    assert (rows, cols) == (orc.rows, orc.cols)
    np.testing.assert_array_equal(wall_bits.reshape(rows, cols), orc.hansen_bits)
    np.testing.assert_array_equal(locs, orc.loc_r[:-1] * cols + orc.loc_c[:-1])
    cells = np.arange(rows * cols)
    for a in range(4):
        o = oracle.TaxiOracle(rows * cols, map=m)
        o.set_state(s=o.encode(cells // cols, cells % cols, 0, 1), elapsed=np.zeros(rows * cols), ndrop=np.zeros(rows * cols))
        o.step(np.full(rows * cols, a))
        r2, c2, _, _ = o.decode(o.s)
        moved = (r2 * cols + c2) != cells
        blocked = (wall_bits >> a) & 1
        ok = ~is_wall                      # the taxi never stands on a wall cell
        np.testing.assert_array_equal(moved[ok], blocked[ok] == 0)


def test_fixed_5x5_wall_bits():
    from gym_po.envs.extended_taxi import parse_taxi_map
    _, _, wb, locs, _ = parse_taxi_map(oracle.TAXI_MAP)
    # SURVEY.md A.1
    assert wb.reshape(5, 5).tolist() == [[5, 9, 5, 1, 9], [4, 8, 4, 0, 8], [4, 0, 0, 0, 8], [12, 4, 8, 4, 8], [14, 6, 10, 6, 10]]
    assert locs.tolist() == [0, 4, 20, 23]


def test_reset_law_is_a_distribution_and_matches_monte_carlo():
    from gym_po.envs.extended_taxi import argmax_multinomial_law, law_to_cdf32
    law = argmax_multinomial_law(500, 300)
    assert abs(law.sum() - 1) < 1e-12 and (law > 0).all() and (np.diff(law) < 0).all()
    rng = np.random.default_rng(0)
    cnt = np.zeros(300)
    for _ in range(8):
        cnt += np.bincount(rng.multinomial(500, np.full(300, 1 / 300), 50_000).argmax(-1), minlength=300)
    chi2 = ((cnt - cnt.sum() * law) ** 2 / (cnt.sum() * law)).sum()
    assert chi2 < 299 + 6 * np.sqrt(2 * 299)
    cdf = law_to_cdf32(law)
    assert cdf.dtype == np.uint32 and (np.diff(cdf.astype(np.int64)) > 0).all() and cdf[-1] == 2**32 - 1
    # tiny case, exact by enumeration: 2 trials over 2 categories -> counts (2,0),(1,1),(0,2) w.p. 1/4,1/2,1/4
    np.testing.assert_allclose(argmax_multinomial_law(2, 2), [0.75, 0.25], atol=1e-12)
    np.testing.assert_allclose(argmax_multinomial_law(3, 2), [0.5, 0.5], atol=1e-12)


def test_msrooms_host_map_matches_oracle():
    """FR_MAP (built from room rectangles) and the multistory walk grid equal the oracle's, which is pinned
    against the real reference; spawn / goal cell lists follow msrooms.py:306-313."""
    from gym_po.envs.rooms import msrooms as M
    from oracle import msrooms as O
    np.testing.assert_array_equal(M.FR_MAP, O.FR_MAP)
    for floors in (1, 2, 5):
        np.testing.assert_array_equal(M.multistory_grid(M.FR_MAP, floors), O.multistory_grid(O.FR_MAP, floors))
    assert tuple(O.UP_YX) == M.UPSTAIRS_YX and tuple(O.DOWN_YX) == M.DOWNSTAIRS_YX
    assert M.END_XYZ == O.END_XYZ


@pytest.mark.parametrize("name", ["render_taxi_5x5", "render_taxi_5x5_hansen", "render_taxi_8x8", "render_taxi_8x8_hansen"])
def test_taxi_render_matches_reference_frames(name):
    """the host-side renderer reproduces, pixel for pixel, the frames the REAL reference drew
    (tests/golden/make_render_golden.py) for the same encoded states and last action."""
    from gym_po.envs.extended_taxi import EXTENDED_TAXI_MAP, TAXI_MAP, parse_taxi_map
    from gym_po.envs.taxi_render import render_taxi
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m = EXTENDED_TAXI_MAP if "8x8" in name else TAXI_MAP
    rows, cols, _, loc_cells, _ = parse_taxi_map(m)
    np_locs = np.concatenate((np.stack(np.divmod(loc_cells, cols), -1), [[-1, -1]]))
    seen = set()
    for s, nm, frame in zip(z["states"], z["names"], z["frames"]):
        got = render_taxi(m, s, len(loc_cells), np_locs, cols, "hansen" in name, str(nm) or None)
        assert got.shape == frame.shape and got.dtype == np.uint8
        np.testing.assert_array_equal(got, frame)
        seen.add(str(nm))
    assert len(seen) >= 2


def test_car_render_matches_reference_frames():
    from gym_po.envs.car_render import render_car
    z = np.load(os.path.join(ROOT, "tests", "golden", "render_car.npz"))
    assert len(set(z["s"][:, 2].tolist())) >= 2                      # with and without the priest indicator
    for s, h, p, frame in zip(z["s"], z["heavens"], z["priests"], z["frames"]):
        np.testing.assert_array_equal(render_car(float(s[0]), float(s[2]), float(h), float(p)), frame)


def test_numa_fallback_parses_nvidia_smi_topo():
    """bind_to_gpu_numa_node's fallback (sysfs numa_node = -1 on virtualised boxes): CPU affinity column of the
    matrix `nvidia-smi topo -m` prints."""
    import os
    from gym_po.sharding import bind_from_topo_text
    mine = sorted(os.sched_getaffinity(0))
    spec = f"{mine[0]}-{mine[-1]}"
    txt = ("\tGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\n"
           f"GPU0\t X \tNV18\tSYS\t{spec}\t0\t\tN/A\nGPU1\tNV18\t X \tSYS\t99990-99999\t1\t\tN/A\nNIC0\tSYS\tSYS\t X \t\t\t\t\n")
    assert bind_from_topo_text(txt, 0, apply=False).startswith("bound to the GPU's CPU affinity")
    assert "no CPU in this process's cpuset" in bind_from_topo_text(txt, 1, apply=False)
    assert "no CPU Affinity column" in bind_from_topo_text("garbage", 0, apply=False)
    assert "not given" in bind_from_topo_text("\tGPU0\tCPU Affinity\nGPU0\t X \t\tN/A\n", 0, apply=False)
