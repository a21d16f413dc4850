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


# ------------------------------------------------------------------ GridPart-compatible decomposer

GRIDPART_REF = os.path.join(REPO, "oracle", "_ref", "GridPart")            # the unmodified reference partitioner
GRIDPART_AUTO = os.path.join(REPO, "oracle", "_ref", "GridPart_auto")      # its commented-out automatic i split restored
GRIDPART_GOLDEN = os.path.join(REPO, "tests", "golden", "gridpart")


def _write_gridin(directory, blocks):
    """blocks: list of dicts(ni, nj, conn=(ixm, ixp, iym, iyp), part=(ipart, jpart), x0, y0) - gridin.conn / gridin.dat as
    GridPart/main.c:23-74 reads them (a smooth, slightly sheared mesh per block)."""
    os.makedirs(os.path.join(directory, "subgrids"), exist_ok=True)
    with open(os.path.join(directory, "gridin.conn"), "w") as f:
        f.write(f"{len(blocks)}\n")
        for b in blocks:
            f.write(" %d %d %d %d   ! ixm ixp iym iyp\n" % b["conn"])
            f.write(" %d %d   ! ipart jpart\n" % b["part"])
    with open(os.path.join(directory, "gridin.dat"), "w") as f:
        for b in blocks:
            f.write(f"{b['ni']} {b['nj']}\n")
            for j in range(b["nj"]):
                for i in range(b["ni"]):
                    x = b["x0"] + 0.01 * i + 0.0007 * j
                    y = b["y0"] + 0.02 * j + 0.003 * np.sin(0.05 * i) * j / b["nj"]
                    f.write("%.16E %.16E \n" % (x, y))


def _tree(directory):
    out = {}
    for name in ["grid.conn"] + sorted(os.path.join("subgrids", n) for n in os.listdir(os.path.join(directory, "subgrids"))):
        with open(os.path.join(directory, name), "rb") as f:
            out[name] = f.read()
    return out


TWO_BLOCKS_MANUAL = [   # stacked in y; 357 nodes in i so that the committed binary's node table {0,99,178,257,356} fits
    dict(ni=357, nj=13, conn=(-1, -3, -4, 1), part=(4, 2), x0=0.0, y0=0.0),
    dict(ni=357, nj=9, conn=(-1, -3, 0, -4), part=(4, 3), x0=0.0, y0=0.24),
]
TWO_BLOCKS_AUTO = [     # side by side in x, uneven splits
    dict(ni=23, nj=11, conn=(-1, 1, -4, -4), part=(3, 2), x0=0.0, y0=0.0),
    dict(ni=14, nj=11, conn=(0, -3, -4, -2), part=(2, 2), x0=0.22, y0=0.0),
]


@pytest.mark.parametrize("which", ["manual", "auto"])
def test_gridpart_output_equals_the_reference_partitioner_bytewise(tmp_path, which):
    """par2d_gridpart + par2d_case_write_gridpart against the reference's own GridPart binary run on the same
    gridin.conn / gridin.dat (a 2-block input: sub-block numbering and conn tuples across source blocks): every
    byte of grid.conn and of every subgrids/grid.block<b>.dat.  'manual': the UNMODIFIED binary, whose i split is the
    hard-coded table istepmanual[] (GridPart/main.c:103-108); 'auto': the automatic split the source keeps in
    comments (sed-restored copy, oracle/build_ref.sh).  The 'auto' output is also pinned by a committed fixture, so
    the test keeps its teeth where oracle/_ref is absent."""
    blocks, binary, isteps = (TWO_BLOCKS_MANUAL, GRIDPART_REF, [0, 99, 178, 257, 356]) if which == "manual" else (TWO_BLOCKS_AUTO, GRIDPART_AUTO, None)
    ref_dir, my_dir = tmp_path / "ref", tmp_path / "mine"
    _write_gridin(str(ref_dir), blocks)
    case = Case.gridpart(str(ref_dir), isteps)
    case.write_gridpart(str(my_dir))
    mine = _tree(str(my_dir))
    assert case.nblocks == sum(b["part"][0] * b["part"][1] for b in blocks)
    if os.path.exists(binary):
        subprocess.run([binary], cwd=str(ref_dir), check=True, capture_output=True)
        ref = _tree(str(ref_dir))
        assert sorted(ref) == sorted(mine)
        for name in ref:
            assert ref[name] == mine[name], name
    elif which == "manual":
        pytest.skip("oracle/_ref/GridPart not built here")
    if which == "auto":
        gold = _tree(GRIDPART_GOLDEN)
        assert sorted(gold) == sorted(mine)
        for name in gold:
            assert gold[name] == mine[name], name


def test_gridpart_reproduces_the_shipped_split_of_the_forward_step(tmp_path):
    """ffstep1 -> 4 x 2: the automatic GridPart rule applied to the single-block forward step gives the shapes, ids,
    conn tuples and BC lists of the reference's shipped Case_ffstep8 (whose files are 1-based and hold 7 digits)."""
    one = Case.load(case_path("Case_ffstep1"))
    x, y = one.nodes(0)
    ib = one.ibound(0)
    ni, nj = one.shape(0)
    codes = (int(ib[-1]), int(ib[ni - 1]), int(ib[0]), int(ib[2 * (ni - 1) + (nj - 1) - 1]))      # left, right, bottom, top
    d = tmp_path / "in"
    os.makedirs(d / "subgrids")
    with open(d / "gridin.conn", "w") as f:
        f.write("1\n %d %d %d %d\n 4 2\n" % tuple(-c for c in codes))
    with open(d / "gridin.dat", "w") as f:
        f.write(f"{ni} {nj}\n")
        for j in range(nj):
            for i in range(ni):
                f.write("%.16E %.16E \n" % (x[j, i], y[j, i] - 1e-4))
    split = Case.gridpart(str(d))
    eight = Case.load(case_path("Case_ffstep8"))
    assert split.nblocks == 8
    for b in range(8):
        assert split.shape(b) == eight.shape(b)
        assert list(split.desc(b).bid) == list(eight.desc(b).bid)
        # (the conn-file BC codes of the shipped case come from an older GridPart - side numbers 0..3, see the
        #  commented-out block_decomp - and are never read by the solver, main.cpp:131-142)
        assert np.array_equal(split.ibound(b), eight.ibound(b))
        xs, ys = split.nodes(b)
        xe, ye = eight.nodes(b)
        assert np.allclose(xs, xe, rtol=0, atol=2e-6) and np.allclose(ys, ye, rtol=0, atol=2e-6)
    if os.path.exists(GRIDPART_AUTO):
        subprocess.run([GRIDPART_AUTO], cwd=str(d), check=True, capture_output=True)
        split.write_gridpart(str(tmp_path / "mine"))
        ref, mine = _tree(str(d)), _tree(str(tmp_path / "mine"))
        for name in ref:
            assert ref[name] == mine[name], name


def test_gridpart_rejects_what_the_reference_rejects(tmp_path):
    bad = [dict(TWO_BLOCKS_AUTO[0]), dict(TWO_BLOCKS_AUTO[1])]
    bad[1]["part"] = (2, 3)                      # jpart mismatch across an x face: GridPart exits with status 100
    _write_gridin(str(tmp_path), bad)
    with pytest.raises(_capi.Par2dError):
        Case.gridpart(str(tmp_path))
    if os.path.exists(GRIDPART_AUTO):
        assert subprocess.run([GRIDPART_AUTO], cwd=str(tmp_path), capture_output=True).returncode == 100
    with pytest.raises(_capi.Par2dError):        # the manual node table does not fit a 23-node block
        _write_gridin(str(tmp_path), TWO_BLOCKS_AUTO)
        Case.gridpart(str(tmp_path), [0, 99, 178, 257, 356])
