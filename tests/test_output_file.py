"""the output / checkpoint file (create_output_file / write_to_file, reference src/io.jl:57-225) written by
libvrt without the HDF5 library, read back by an independent reader that is itself pinned against a genuine libhdf5 file.
CPU only (file writing needs no device)."""
import os

import numpy as np
import pytest

import h5mini_reader as H
from conftest import GOLDEN


def test_reader_reads_a_file_written_by_libhdf5():
    """the reader's own pin: a MATLAB v7.3 file (512-byte user block + a file written by the HDF5 library)"""
    f = H.File(os.path.join(GOLDEN, "libhdf5_sample.mat"))
    assert f.keys() == ["testdouble"] and (f.leaf_k, f.internal_k) == (4, 16)
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
    assert np.allclose(d.read()[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)


VORONOI = {"source_function": "nlam,n", "populations": "n,3", "positions": "3,n", "temperature": "n", "hydrogen_populations": "n",
           "electron_density": "n", "velocity_z": "n", "velocity_x": "n", "velocity_y": "n", "boundaries": "6", "convergence": "it",
           "n_bb": "1", "n_bf": "1", "wavelength": "nlam", "line_center": "1", "time": "1"}


def test_voronoi_output_file_round_trip(tmp_path):
    from voronoirt_b200 import api
    n, nlam, maxiter = 1234, 19, 150
    rng = np.random.default_rng(5)
    path = tmp_path / "out.h5"
    out = api.create_output_file(path, nlam, n, maxiter)
    S = np.asfortranarray(rng.random((nlam, n)))
    pops = np.asfortranarray(rng.random((n, 3)))
    pos = np.asfortranarray(rng.random((3, n)))
    T = rng.random(n)
    out.write("source_function", S)
    out.write("populations", pops)
    out.write("positions", pos)
    out.write("temperature", T)
    out.write("boundaries", np.arange(6.0))
    api.write_to_file(0.25, out, 1)
    api.write_to_file(0.125, out, 2)
    api.write_to_file(50, out, "n_bb")
    api.write_to_file(20, out, "n_bf")
    out.write("wavelength", np.linspace(90, 130, nlam))
    out.write("line_center", np.array([121.5]))
    out.write("time", np.array([3.5]))
    with pytest.raises(Exception):
        out.write("temperature", np.zeros(n + 1))       # wrong size
    with pytest.raises(Exception):
        out.write("no_such_dataset", np.zeros(1))
    out.close()

    f = H.File(str(path))
    assert f.keys() == sorted(VORONOI)                    # names of io.jl:196-225
    dims = {"n": n, "nlam": nlam, "it": maxiter + 1}
    for name, shape in VORONOI.items():
        julia = tuple(dims.get(t, int(t) if t.isdigit() else None) for t in shape.split(","))
        d = f[name]
        assert d.shape == julia[::-1], name              # HDF5.jl stores the reversed shape over the same bytes
        assert d.dtype == (np.dtype("<i8") if name in ("n_bb", "n_bf") else np.dtype("<f8")), name
        assert d.layout[0] == "contiguous"
    # a reader of the reference (h5py: C order of the reversed shape; HDF5.jl: column-major Julia shape) sees the same memory
    assert np.array_equal(f["source_function"].read(), S.T)
    assert np.array_equal(f["populations"].read(), pops.T)
    assert np.array_equal(f["positions"].read(), pos.T)
    assert np.array_equal(f["temperature"].read(), T)
    conv = f["convergence"].read()
    assert conv[0] == 0.25 and conv[1] == 0.125 and not conv[2:].any()      # created as zeros (io.jl:216)
    assert f["n_bb"].read()[0] == 50 and f["n_bf"].read()[0] == 20
    assert f["line_center"].read()[0] == 121.5 and f["time"].read()[0] == 3.5
    assert f.eof == os.path.getsize(path)


def test_regular_output_file_shapes(tmp_path):
    from voronoirt_b200 import api
    nz, nx, ny, nlam = 7, 5, 6, 4
    path = tmp_path / "reg.h5"
    out = api.create_output_file(path, nlam, (nz, nx, ny), 10)
    S = np.asfortranarray(np.random.default_rng(1).random((nlam, nz, nx, ny)))
    out.write("source_function", S)
    out.write("z", np.arange(nz, dtype=float))
    out.close()
    f = H.File(str(path))
    assert set(f.keys()) == {"source_function", "populations", "z", "x", "y", "temperature", "hydrogen_populations", "electron_density",
                             "velocity_z", "velocity_x", "velocity_y", "convergence", "n_bb", "n_bf", "wavelength", "line_center", "time"}
    assert f["source_function"].shape == (ny, nx, nz, nlam) and f["populations"].shape == (3, ny, nx, nz)
    assert np.array_equal(f["source_function"].read(), S.transpose(3, 2, 1, 0))
    assert np.array_equal(f["z"].read(), np.arange(nz))


@pytest.mark.gpu
def test_write_state_from_the_device(tmp_path):
    """vrt_output_write_state: S and populations go from the solver's device state into the file in host site order"""
    import voronoirt_b200 as V
    from conftest import load_grid
    from voronoirt_b200 import api, synth
    pos, nbr, b = load_grid("grid_strat3000")
    n = pos.shape[1]
    a = synth.atmosphere(pos[0], pos[1], pos[2])
    cell = V.read_cell(nbr, n, pos, b[2], b[3], b[4], b[5])
    sites = V.VoronoiSites(*cell, a["temperature"], a["electron_density"], a["hydrogen_density"], a["velocity_z"], a["velocity_x"],
                           a["velocity_y"], b[0], b[1], b[2], b[3], b[4], b[5], n)
    line, lte, α_cont, ελ, Cr = synth.line_inputs(a["temperature"], a["electron_density"], a["hydrogen_density"], 10, 4)
    solver = V.Solver(sites, V.quadrature_path("ul2n3"), line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte)
    path = tmp_path / "state.h5"
    out = api.create_output_file(path, len(line.λ), n, 5)
    diffs = []

    def cb(rec):
        out.write_state(solver)
        out.write_convergence(rec["iteration"], rec["diff"])
        diffs.append(rec["diff"])
    solver.iterate(-1.0, 2, cb)
    api.write_to_file(sites, out)
    api.write_to_file(line, out)
    S, J, pops = solver.get_state()
    out.close()
    solver.close()
    f = H.File(str(path))
    assert np.array_equal(f["source_function"].read(), S.T)
    assert np.array_equal(f["populations"].read(), pops.T)
    assert np.array_equal(f["positions"].read(), pos.T)
    assert np.array_equal(f["convergence"].read()[:2], np.array(diffs))
    assert np.array_equal(f["wavelength"].read(), line.λ)


def test_messages_are_byte_identical_to_what_libhdf5_writes(tmp_path):
    """the datatype and fill-value messages of a Float64 dataset, byte for byte, against the same messages in the file written by
    the HDF5 library; dataspace and symbol-table structures of both files pass the same parser (above)"""
    from voronoirt_b200 import api
    path = tmp_path / "m.h5"
    api.create_output_file(path, 3, 5, 2).close()
    ours, real = H.File(str(path)), H.File(os.path.join(GOLDEN, "libhdf5_sample.mat"))

    def messages(f, name):
        return {t: (fl, bytes(d)) for t, fl, d in H._messages(f, f.root.entries[name][0])}
    a, b = messages(ours, "temperature"), messages(real, "testdouble")
    assert a[0x0003] == b[0x0003]                 # IEEE binary64 little-endian datatype, flags included
    assert a[0x0005] == b[0x0005]                 # fill value message (version 1: late allocation, written if set, default value)
    assert a[0x0001][1][:8] == bytes([1, 1, 0, 0, 0, 0, 0, 0]) and b[0x0001][1][:8] == bytes([1, 2, 0, 0, 0, 0, 0, 0])   # dataspace v1, rank 1 / 2
    # superblock fields other than addresses and K values
    so, sr = ours.buf[:16], real.buf[512:528]
    assert so == sr
