"""sfm_data.bin / matches.*.bin: the cereal PortableBinary containers of the reference, read and written
without cereal / OpenMVG.  PARITY UNPINNED (no reference-written file exists and none can be produced in this image):
the writer is pinned against a byte stream assembled HERE, independently, field by field from the published
serialisation code (cereal portable_binary / memory / polymorphic / map / vector / string; OpenMVG sfm_data_io_cereal,
sfm_view, sfm_view_priors, Camera_Pinhole*, pose3, sfm_landmark) -- the same information a maintainer would check a real
file against."""
import struct

import numpy as np

MSB, MSB2 = 0x80000000, 0x40000000


def u8(v): return struct.pack("<B", v)
def u32(v): return struct.pack("<I", v)
def u64(v): return struct.pack("<Q", v)
def f64(v): return struct.pack("<d", v)
def string(s): return u64(len(s)) + s.encode()
def vec(v): return u64(len(v)) + b"".join(f64(x) for x in v)


def _fill(sd, r3dlib):
    sd.root_path = "pictures/set0"
    sd.add_view(0, "image000000.jpg", 1920, 1080, id_intrinsic=0, id_pose=0)
    sd.add_view(1, "image000001.jpg", 1920, 1080, id_intrinsic=0, id_pose=1, prior_center=(4.0e6, 3.0e5, 4.9e6))
    sd.add_view(2, "image000002.jpg", 640, 480, id_intrinsic=1, id_pose=2, local_path="sub")
    sd.add_intrinsic(0, r3dlib.CAM_RADIAL3, 1920, 1080, 2112.0, 960.0, 540.0, (0.01, -0.02, 0.003))
    sd.add_intrinsic(1, r3dlib.CAM_PINHOLE, 640, 480, 704.0, 320.0, 240.0)
    return sd


def _expected_views_intrinsics():
    b = u8(1) + string("0.3") + string("pictures/set0")
    # views (std::map order): shared_ptr<View>
    b += u64(3)
    b += u32(0) + u32(MSB2) + u32(1 | MSB) + string("") + string("image000000.jpg") + u32(1920) + u32(1080) + u32(0) + u32(0) + u32(0)
    b += (u32(1) + u32(1 | MSB) + string("view_priors") + u32(2 | MSB) + string("") + string("image000001.jpg") + u32(1920) + u32(1080)
          + u32(1) + u32(0) + u32(1) + u8(1) + vec([1.0, 1.0, 1.0]) + vec([4.0e6, 3.0e5, 4.9e6]))
    b += u32(2) + u32(MSB2) + u32(3 | MSB) + string("sub") + string("image000002.jpg") + u32(640) + u32(480) + u32(2) + u32(1) + u32(2)
    # intrinsics: shared_ptr<IntrinsicBase>: polymorphic ids continue after "view_priors" (= 1), pointer ids after the views
    b += u64(2)
    b += (u32(0) + u32(2 | MSB) + string("pinhole_radial_k3") + u32(4 | MSB) + u32(1920) + u32(1080) + f64(2112.0) + vec([960.0, 540.0])
          + vec([0.01, -0.02, 0.003]))
    b += u32(1) + u32(3 | MSB) + string("pinhole") + u32(5 | MSB) + u32(640) + u32(480) + f64(704.0) + vec([320.0, 240.0])
    return b


def test_views_and_intrinsics_byte_stream(r3dlib, tmp_path):
    """What R3DProject::writeSfmData stores (VIEWS | INTRINSICS, src/R3DProject.cpp:1298-1302)."""
    sd = _fill(r3dlib.SfmData(), r3dlib)
    p = tmp_path / "sfm_data.bin"
    sd.save(str(p), r3dlib.SFM_VIEWS | r3dlib.SFM_INTRINSICS)
    want = _expected_views_intrinsics() + u64(0) + u64(0) + u64(0)      # empty extrinsics, structure, control points
    assert p.read_bytes() == want
    back = r3dlib.SfmData.load(str(p))
    assert back.root_path == "pictures/set0"
    v = back.views()
    assert [x["filename"] for x in v] == ["image000000.jpg", "image000001.jpg", "image000002.jpg"]
    assert v[1]["has_prior"] and v[1]["pose_center"] == [4.0e6, 3.0e5, 4.9e6] and not v[0]["has_prior"]
    assert v[2]["local_path"] == "sub" and v[2]["id_intrinsic"] == 1
    i = back.intrinsics()
    assert i[0]["model"] == r3dlib.CAM_RADIAL3 and i[0]["disto"][:3] == [0.01, -0.02, 0.003] and i[1]["model"] == r3dlib.CAM_PINHOLE


def test_full_sfm_data_roundtrip_and_byte_stream(r3dlib, tmp_path):
    """After SfM the reference saves ALL (src/threads/R3DTriangulationThread.cpp:453-455): + poses and landmarks."""
    sd = _fill(r3dlib.SfmData(), r3dlib)
    R = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    sd.add_pose(0, np.eye(3), (0.0, 0.0, 0.0))
    sd.add_pose(2, R, (1.5, -2.0, 0.25))
    sd.add_landmark(7, (1.0, 2.0, 3.0), [(0, 11, 100.5, 200.25), (2, 5, 10.0, 20.0)])
    sd.add_landmark(9, (-1.0, 0.5, 8.0), [(1, 3, 7.0, 8.0), (0, 4, 1.0, 2.0)])
    sd.add_landmark(1, (0.0, 0.0, 1.0), [(0, 0, 5.0, 6.0)], control_point=True)
    for model, disto in ((r3dlib.CAM_RADIAL1, (0.1,)), (r3dlib.CAM_BROWN, (0.1, 0.2, 0.3, 0.01, 0.02)), (r3dlib.CAM_FISHEYE, (0.1, 0.2, 0.3, 0.4))):
        sd.add_intrinsic(10 + model, model, 100, 80, 90.0, 50.0, 40.0, disto)
    p = tmp_path / "sfm_data.bin"
    sd.save(str(p), r3dlib.SFM_ALL)
    raw = p.read_bytes()
    tail = u64(2)                                                        # extrinsics
    tail += u32(0) + u64(3) + vec([1.0, 0.0, 0.0]) + vec([0.0, 1.0, 0.0]) + vec([0.0, 0.0, 1.0]) + vec([0.0, 0.0, 0.0])
    tail += u32(2) + u64(3) + vec(R[0]) + vec(R[1]) + vec(R[2]) + vec([1.5, -2.0, 0.25])
    tail += u64(2)                                                       # structure: Landmark{X, observations map{view: {id_feat, x}}}
    tail += u32(7) + vec([1.0, 2.0, 3.0]) + u64(2) + u32(0) + u32(11) + vec([100.5, 200.25]) + u32(2) + u32(5) + vec([10.0, 20.0])
    tail += u32(9) + vec([-1.0, 0.5, 8.0]) + u64(2) + u32(0) + u32(4) + vec([1.0, 2.0]) + u32(1) + u32(3) + vec([7.0, 8.0])
    tail += u64(1) + u32(1) + vec([0.0, 0.0, 1.0]) + u64(1) + u32(0) + u32(0) + vec([5.0, 6.0])   # control points
    assert raw.endswith(tail)
    assert b"pinhole_radial_k1" in raw and b"pinhole_brown_t2" in raw and b"pinhole_fisheye" in raw
    back = r3dlib.SfmData.load(str(p))
    p2 = tmp_path / "again.bin"
    back.save(str(p2), r3dlib.SFM_ALL)
    assert p2.read_bytes() == raw                                        # load -> save is the identity
    assert [x["id"] for x in back.poses()] == [0, 2] and np.allclose(back.poses()[1]["R"], R)
    lm = back.landmarks()
    assert [x["id"] for x in lm] == [7, 9] and lm[1]["obs"] == [(0, 4, 1.0, 2.0), (1, 3, 7.0, 8.0)]
    assert len(back.landmarks(control_points=True)) == 1
    assert {x["model"] for x in back.intrinsics()} == {1, 2, 3, 4, 5}


def test_truncated_and_foreign_files_fail_cleanly(r3dlib, tmp_path):
    sd = _fill(r3dlib.SfmData(), r3dlib)
    p = tmp_path / "sfm_data.bin"
    sd.save(str(p))
    raw = p.read_bytes()
    for cut in (0, 1, 10, len(raw) // 2, len(raw) - 3):
        q = tmp_path / ("cut%d.bin" % cut)
        q.write_bytes(raw[:cut])
        try:
            r3dlib.SfmData.load(str(q))
        except r3dlib.R3DError:
            continue
        raise AssertionError("truncated file (%d bytes) loaded" % cut)
    q = tmp_path / "big_endian.bin"
    q.write_bytes(b"\x00" + raw[1:])
    try:
        r3dlib.SfmData.load(str(q))
    except r3dlib.R3DError as e:
        assert e.code == -5
    else:
        raise AssertionError("big-endian archive accepted")


def test_matches_bin_format_and_roundtrip(r3dlib, tmp_path):
    pairs = np.array([[0, 2], [0, 1], [3, 4]], np.uint32)
    ofs = np.array([0, 2, 3, 5], np.uint64)
    m = np.array([(1, 1), (2, 2), (7, 7), (8, 8), (9, 9)], r3dlib.indmatch_dtype)
    mm = r3dlib.Matches.from_csr(pairs, ofs, m)
    p = tmp_path / "matches.f.bin"
    mm.save(str(p))
    want = u8(1) + u64(3)
    want += u32(0) + u32(1) + u64(1) + u32(7) + u32(7)
    want += u32(0) + u32(2) + u64(2) + u32(1) + u32(1) + u32(2) + u32(2)
    want += u32(3) + u32(4) + u64(2) + u32(8) + u32(8) + u32(9) + u32(9)
    assert p.read_bytes() == want
    back = r3dlib.Matches.load(str(p))
    d = back.to_dict()
    assert sorted(d) == [(0, 1), (0, 2), (3, 4)] and d[(3, 4)]["j"].tolist() == [8, 9]
    t = tmp_path / "matches.f.txt"
    back.save(str(t))
    assert r3dlib.Matches.load(str(t)).to_dict().keys() == d.keys()
