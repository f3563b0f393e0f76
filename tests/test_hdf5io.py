"""hdf5io.py: the HDF5 subset behind `-f file.h5` (GCloudDmriSolver.py:150-177) and the pre-processing output
(PreprocessingMultiCompt.py:148-152).  No libhdf5 here, so the reader is exercised (i) on the writer's output
(superblock v0, v1 object headers, symbol-table groups spanning several symbol-table nodes, contiguous datasets,
attributes) and (ii) on a hand-built file in the OTHER dialect libhdf5 can produce (superblock v2, v2 object headers,
link-message groups, compact and chunked + deflate + shuffle datasets), assembled here byte by byte from the format
specification."""
import struct
import zlib

import numpy as np
import pytest

from dmri_fem_cloud_b200 import hdf5io, meshes, preprocess


def test_dolfin_container_round_trip(tmp_path):
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 2.0, (2, 1, 1), 10, 1)
    nc = len(tets)
    rng = np.random.default_rng(0)
    fields = {"phase": (marker % 2).astype(float), "T2": rng.uniform(1e4, 1e6, nc), "ic": np.ones(nc)}
    for a in range(3):
        for b in range(3):
            fields["d%d%d" % (a, b)] = rng.uniform(1e-3, 3e-3, nc) if a == b else np.zeros(nc)
    path = str(tmp_path / "files.h5")
    hdf5io.write_dolfin_h5(path, xyz, tets, fields)
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and struct.unpack_from("<Q", raw, 40)[0] == len(raw)      # end-of-file address
    f = hdf5io.File(path)
    assert f.keys("") == sorted(["mesh"] + list(fields))          # 13 members: two symbol-table nodes under one B-tree
    assert f.keys("mesh") == ["coordinates", "topology"]
    assert f.attrs("mesh/topology")["celltype"] == "tetrahedron"
    assert f.attrs("T2")["signature"] == "FiniteElement('Discontinuous Lagrange', tetrahedron, 0)"
    assert f["mesh/topology"].dtype == np.int64 and f["T2/x_cell_dofs"].dtype == np.uint64
    got = hdf5io.read_dolfin_h5(path)
    assert np.array_equal(got["xyz"], xyz) and np.array_equal(got["tets"], tets)
    for k, v in fields.items():
        assert np.array_equal(got[k], v), k
    with pytest.raises(KeyError):
        f["mesh/nothing"]
    with pytest.raises(hdf5io.HDF5Error):
        open(tmp_path / "bad.h5", "wb").write(b"not hdf5" * 100)
        hdf5io.File(str(tmp_path / "bad.h5"))


def test_dg0_dofs_in_another_order_and_triangles(tmp_path):
    """DOLFIN numbers DG0 dofs and lists cells in its own order: value of cell cells[i] = vector[cell_dofs[x[i]]]."""
    xy = np.array([[0, 0], [1, 0], [1, 1], [0, 1], [2, 0.5]], dtype=float)
    tris = np.array([[0, 1, 2], [0, 2, 3], [1, 4, 2]], dtype=np.int32)
    w = hdf5io.Writer()
    w.dataset("mesh/coordinates", xy)
    w.dataset("mesh/topology", tris.astype(np.int64), attrs={"celltype": "triangle"})
    w.dataset("phase/vector_0", np.array([30.0, 10.0, 20.0]))          # dof 0 -> cell 2, dof 1 -> cell 0, dof 2 -> cell 1
    w.dataset("phase/cell_dofs", np.array([0, 2, 1], dtype=np.int32))  # listed for cells [2, 1, 0]
    w.dataset("phase/x_cell_dofs", np.array([0, 1, 2, 3], dtype=np.uint64))
    w.dataset("phase/cells", np.array([2, 1, 0], dtype=np.uint64))
    w.save(str(tmp_path / "t.h5"))
    got = hdf5io.read_dolfin_h5(str(tmp_path / "t.h5"))
    assert got["tets"].shape == (3, 3) and got["xyz"].shape == (5, 2)
    assert np.array_equal(got["phase"], [10.0, 20.0, 30.0])


def test_preprocess_writes_h5_like_the_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 2.0, (2, 1, 1), 8, 1)
    monkeypatch.setattr(preprocess, "_load", lambda p: (xyz, tets))
    monkeypatch.setattr(preprocess.meshes, "read_dolfin_markers", lambda p: marker)
    assert preprocess.main(["preprocess", "-m", "mesh.xml", "-pmk", "pmk.xml", "-D0", "3e-3", "1e-3", "3e-3", "-o", "out.xml"]) == 0
    got = hdf5io.read_dolfin_h5("out.h5")          # filename + '.h5' (PreprocessingMultiCompt.py:144-146)
    assert set(got) == {"xyz", "tets", "phase", "T2", "ic"} | {"d%d%d" % (a, b) for a in range(3) for b in range(3)}
    assert np.array_equal(got["phase"], marker % 2) and np.array_equal(got["d11"], np.array([3e-3, 1e-3, 3e-3])[marker])
    assert np.all(got["T2"] == 1e6) and np.all(got["d01"] == 0)


# ---------------------------------------------------------------------------------------------------------------
# the other dialect, built by hand: superblock v2, OHDR v2, link messages, compact + chunked/deflate/shuffle data

def _msg2(mtype, data):
    return struct.pack("<BHB", mtype, len(data), 0) + data


def _ohdr2(msgs):
    body = b"".join(msgs)
    return b"OHDR" + struct.pack("<BB", 2, 0x01) + struct.pack("<H", len(body)) + body + b"\0\0\0\0"      # 2-byte chunk size, checksum unchecked


def _link(name, addr):
    return _msg2(0x06, struct.pack("<BB", 1, 0) + struct.pack("<B", len(name)) + name.encode() + struct.pack("<Q", addr))


def test_reader_on_new_style_file_with_chunked_deflate(tmp_path):
    f64 = struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    i32 = struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, 4, 0, 32)
    out = bytearray(48)

    def alloc(b):
        while len(out) % 8:
            out.append(0)
        a = len(out)
        out.extend(b)
        return a

    # compact int32 dataset (3,)
    small = np.array([7, -8, 9], dtype="<i4")
    d_small = alloc(_ohdr2([_msg2(0x01, struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 3)), _msg2(0x03, i32),
                            _msg2(0x08, struct.pack("<BBH", 3, 0, small.nbytes) + small.tobytes())]))
    # chunked float64 dataset (5, 3), chunks (2, 3), shuffle + deflate
    big = np.arange(15, dtype="<f8").reshape(5, 3) * 1.5
    keys = []
    for r0 in range(0, 5, 2):
        chunk = np.zeros((2, 3), dtype="<f8")
        chunk[:min(2, 5 - r0)] = big[r0:r0 + 2]
        shuf = np.frombuffer(chunk.tobytes(), np.uint8).reshape(-1, 8).T.tobytes()
        comp = zlib.compress(shuf)
        keys.append((len(comp), (r0, 0, 0), alloc(comp)))
    bt = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), hdf5io.UNDEF, hdf5io.UNDEF)
    for size, offs, addr in keys:
        bt += struct.pack("<II3Q", size, 0, *offs) + struct.pack("<Q", addr)
    bt += struct.pack("<II3Q", 0, 0, 6, 0, 0)
    btree = alloc(bt)
    pipeline = struct.pack("<BB", 2, 2) + struct.pack("<HHH", 2, 0, 1) + struct.pack("<I", 8) + \
        struct.pack("<HHH", 1, 0, 1) + struct.pack("<I", 6)
    d_big = alloc(_ohdr2([_msg2(0x01, struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 5, 3)), _msg2(0x03, f64),
                          _msg2(0x0B, pipeline),
                          _msg2(0x08, struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", btree) + struct.pack("<III", 2, 3, 8))]))
    grp = alloc(_ohdr2([_link("small", d_small), _link("big", d_big)]))
    root = alloc(_ohdr2([_link("g", grp)]))
    out[0:48] = hdf5io.SIGNATURE + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, hdf5io.UNDEF, len(out), root) + b"\0" * 4
    path = str(tmp_path / "new.h5")
    open(path, "wb").write(bytes(out))
    f = hdf5io.File(path)
    assert f.keys("") == ["g"] and f.keys("g") == ["big", "small"]
    assert np.array_equal(f["g/small"], small)
    assert np.array_equal(f["g/big"], big)
