"""Host logic above the C-ABI seam (no GPU): case directories, numbering bases, GridPart-style
splitting, partitioning, writers, and the shape of the C ABI itself."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import CASES, REPO, THERM, case_path, make_cfg
from par2d_b200 import Case, Solver, _capi, default_config, load_thermo, write_solution
from par2d_b200._capi import BOTTOM, LEFT, RIGHT, TOP


def test_case_loader_shapes_and_bases():
    one = Case.load(case_path("Case_ffstep1"))          # 1-based file name, single block
    assert one.nblocks == 1 and one.shape(0) == (398, 140) and one.ncells == 55183
    assert list(one.desc(0).bid) == [-1, -1, -1, -1]
    ib = one.ibound(0)
    assert len(ib) == 2 * 397 + 2 * 139
    assert set(ib[:397]) == {4} and set(ib[397:397 + 139]) == {3} and set(ib[-139:]) == {1}
    eight = Case.load(case_path("Case_ffstep8"))        # 1-based ids -> shifted to 0-based
    assert eight.nblocks == 8 and eight.ncells == 55183
    assert [eight.shape(b)[0] for b in range(4)] == [101, 101, 101, 98]
    d0, d5 = eight.desc(0), eight.desc(5)
    assert d0.bid[RIGHT] == 1 and d0.bid[TOP] == 4 and d0.bid[LEFT] == -1 and d0.bid[BOTTOM] == -1
    assert d5.bid[LEFT] == 4 and d5.bid[RIGHT] == 6 and d5.bid[BOTTOM] == 1 and d5.bid[TOP] == -1
    noz = Case.load(case_path("Case_rktnoz9"))          # 0-based
    assert noz.nblocks == 9 and noz.ncells == 42842 and noz.shape(4) == (20, 11)
    assert noz.desc(3).bid[BOTTOM] == 4 and noz.desc(4).bid[TOP] == 3
    spk = Case.load(case_path("Case_spikefine24"))
    assert spk.nblocks == 24 and spk.ncells == 49000
    assert spk.desc(4).bid[LEFT] == 19 and spk.desc(19).bid[RIGHT] == 4


def test_y_shift_applied_on_load():
    """FileIO.cpp:41 adds 1e-4 to every y."""
    c = Case.load(case_path("Case_ffstep8"))
    _, y = c.nodes(0)
    with open(os.path.join(case_path("Case_ffstep8"), "subgrids", "grid.block1.dat")) as f:
        lines = f.read().splitlines()
    y_file = float(lines[3 + 101 * 5 + 7].split()[1])
    assert y[5, 7] == y_file + 1e-4


def test_split_follows_gridpart_slicing():
    """ffstep8 is GridPart's 4x2 split of ffstep1's index space."""
    one = Case.load(case_path("Case_ffstep1"))
    split = one.split(4, 2)
    eight = Case.load(case_path("Case_ffstep8"))
    assert split.nblocks == 8
    for b in range(8):
        assert split.shape(b) == eight.shape(b)
        assert list(split.desc(b).bid) == list(eight.desc(b).bid)
        xs, ys = split.nodes(b)
        xe, ye = eight.nodes(b)
        assert np.allclose(xs, xe, rtol=0, atol=2e-6) and np.allclose(ys, ye, rtol=0, atol=2e-6)   # %le files: 7 digits
        # exterior faces carry the original codes, interior faces -1
        se, ee = split.ibound(b), eight.ibound(b)
        assert np.array_equal(se, ee)


def test_case_write_roundtrip(tmp_path):
    c = Case.synthetic_ramp(48, 32, 3, 2)
    c.write(str(tmp_path / "zero"), one_based=False)
    c.write(str(tmp_path / "one"), one_based=True)
    a, b = Case.load(str(tmp_path / "zero")), Case.load(str(tmp_path / "one"))
    assert a.nblocks == b.nblocks == 6
    for k in range(6):
        assert list(a.desc(k).bid) == list(b.desc(k).bid) == list(c.desc(k).bid)
        assert np.array_equal(a.ibound(k), c.ibound(k))
        xa, ya = a.nodes(k)
        xc, yc = c.nodes(k)
        assert np.array_equal(xa, xc) and np.allclose(ya, yc, rtol=0, atol=1e-15)
        assert np.array_equal(xa, b.nodes(k)[0])


def test_synthetic_ramp_geometry():
    c = Case.synthetic_ramp(64, 32, 4, 2)
    assert c.nblocks == 8 and c.ncells == 64 * 32
    x, y = c.nodes(0)
    assert x[0, 0] == 0.0 and y[0, 0] == 1e-4
    x7, y7 = c.nodes(7)
    assert x7[-1, -1] == 3.0 and abs(y7[-1, -1] - (1.0 + 1e-4)) < 1e-15
    # neighbours share their node line exactly
    x1, y1 = c.nodes(1)
    assert np.array_equal(x[:, -1], x1[:, 0]) and np.array_equal(y[:, -1], y1[:, 0])
    assert set(c.ibound(0)[:16]) == {4} and set(c.ibound(0)[16:32]) == {-1} and set(c.ibound(0)[-16:]) == {1}


def test_partition_patches_and_lpt():
    c = Case.synthetic_ramp(64, 64, 8, 8)
    own = c.partition(8)
    assert sorted(set(own)) == list(range(8)) and all(own.count(r) == 8 for r in range(8))
    # contiguous 2x4 / 4x2 patches: every rank's blocks form a rectangle
    for r in range(8):
        ids = [b for b in range(64) if own[b] == r]
        ii, jj = sorted({b % 8 for b in ids}), sorted({b // 8 for b in ids})
        assert len(ii) * len(jj) == 8
    noz = Case.load(case_path("Case_rktnoz9"))
    own = noz.partition(8)
    loads = [sum((noz.shape(b)[0] - 1) * (noz.shape(b)[1] - 1) for b in range(9) if own[b] == r) for r in range(8)]
    assert max(loads) == 14651 and min(loads) > 0     # the two 14651-cell blocks bound the balance
    assert noz.partition(1) == [0] * 9


def test_thermo_file_matches_default_config(oracle):
    cfg = default_config(oracle)
    loaded = load_thermo(THERM, default_config(oracle))
    assert list(loaded.thermo_hi) == list(cfg.thermo_hi) and list(loaded.thermo_lo) == list(cfg.thermo_lo)
    assert loaded.Rs == cfg.Rs == 8314.34 / 28.9


def test_solution_writer_format(tmp_path, oracle):
    """print_state byte format (FileIO.cpp:84-115), checked against the reference's own output when available."""
    case = Case.load(case_path("Case_rktnoz9"))
    s = Solver(case, make_cfg(oracle), lib=oracle)
    s.step(3)
    x, y = case.nodes(4)
    path = tmp_path / "solution.b5.tec"
    write_solution(str(path), "Final State", s.cfg, x, y, s.get_state(4))
    text = path.read_text().splitlines()
    assert text[0] == 'TITLE = "Final State"' and text[2] == "ZONE I=19, J=10, DATAPACKING=POINT"
    assert len(text) == 3 + 19 * 10
    assert re.fullmatch(r"(-?\d\.\d{16}e[+-]\d\d,\t ){8}-?\d\.\d{16}e[+-]\d\d ", text[3])
    import run_ref
    if run_ref.available():
        shapes = [case.shape(b) for b in range(9)]
        out = run_ref.run(case_path("Case_rktnoz9"), THERM, shapes, 3, keep=True)
        ref_text = open(os.path.join(out["run_dir"], "Outputs", "solution.b5.tec")).read().splitlines()
        assert text == ref_text


def _declared(header):
    src = open(os.path.join(REPO, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(par2d_[a-z0-9_]+)\s*\(", src)))


def test_c_abi_libraries_export_every_declared_symbol():
    """No compute calls here: the CUDA library must load without a GPU and export the whole header."""
    gpu_syms = _declared("par2d_gpu.h")
    host_syms = _declared("par2d_host.h")
    assert len(gpu_syms) >= 25 and len(host_syms) >= 10
    gpu = C.CDLL(_capi.GPU_LIB)
    host = C.CDLL(_capi.HOST_LIB)
    for s in gpu_syms:
        assert hasattr(gpu, s), f"libpar2d_gpu.so lacks {s}"
    for s in host_syms:
        assert hasattr(host, s), f"libpar2d_host.so lacks {s}"
    # the product library does not link the oracle
    deps = subprocess.run(["ldd", _capi.GPU_LIB], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_struct_layouts_match_the_header():
    assert C.sizeof(_capi.Config) == 8 + 8 * 3 + 32 + 16 + 56 * 2 + 16
    assert C.sizeof(_capi.BlockDesc) == 8 + 24 + 32 + 8
    assert C.sizeof(_capi.StepInfo) == 48


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    gpu = _capi.load_gpu()
    case = Case.load(case_path("Case_rktnoz9"))
    with pytest.raises(_capi.Par2dError) as e:
        Solver(case, default_config(gpu), lib=gpu)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_create_rejects_inconsistent_tables(oracle):
    case = Case.load(case_path("Case_ffstep8"))
    descs = case.descs()
    descs[0].bid[RIGHT] = 5          # block 5 does not point back at block 0
    h = C.c_void_p()
    cfg = default_config(oracle)
    assert oracle.create(C.byref(cfg), 8, descs, C.byref(h)) == 1


def test_solution_reader_round_trips_the_writer(tmp_path):
    """par2d_read_solution (restart reader) inverts par2d_write_solution / print_state
    (FV/FileIO.cpp:84-115) exactly: %.16e round-trips doubles; a wrong zone size is refused."""
    from par2d_b200.host import read_solution, write_solution
    from par2d_b200._capi import Par2dError
    rng = np.random.default_rng(7)
    case = Case.load(case_path("Case_spikefine24"))
    x, y = case.nodes(3)
    ny, nx = x.shape
    unk = rng.uniform(0.1, 3000.0, size=(ny - 1, nx - 1, 4))
    unk[..., 3] = rng.uniform(250.0, 3000.0, size=(ny - 1, nx - 1))    # T above the 201 K clamp the writer applies
    cfg = make_cfg(_capi.load_oracle())
    path = str(tmp_path / "solution.b4.tec")
    write_solution(path, "Final State", cfg, x, y, unk)
    back, xc, yc = read_solution(path, nx, ny)
    assert np.array_equal(back, unk)
    assert np.array_equal(xc, 0.25 * (x[:-1, :-1] + x[:-1, 1:] + x[1:, 1:] + x[1:, :-1]))
    with pytest.raises(Par2dError):
        read_solution(path, nx + 1, ny)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's own CPU implementation on the host cores, bounded sample): one JSON
    line with the driver's keys.  Needs oracle/_ref (the unmodified reference build) or falls back to the oracle port."""
    import json
    import sys
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "cell-updates/s" and line["value"] > 1e5
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 3
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
