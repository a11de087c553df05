"""CPU-only: PLY reader/writer and surface sampler of the headless shape-transfer harness."""
import numpy as np

from deformationpyramid_b200 import shape_transfer as st


def _tetra(tmp_path):
    p = tmp_path / "t.ply"
    p.write_text("ply\nformat ascii 1.0\ncomment c\nelement vertex 4\nproperty float x\nproperty float y\n"
                 "property float z\nproperty float nx\nproperty float ny\nproperty float nz\nproperty float s\n"
                 "property float t\nelement face 3\nproperty list uchar uint vertex_indices\nend_header\n"
                 "0 0 0 0 0 1 0 0\n1 0 0 0 0 1 0 0\n0 1 0 0 0 1 0 0\n0 0 1 0 0 1 0 0\n"
                 "3 0 1 2\n3 0 1 3\n4 0 2 3 1\n")
    return p


def test_ply_roundtrip_and_polygon_fan(tmp_path):
    v, f, props = st.read_ply_ascii(str(_tetra(tmp_path)))
    assert v.shape == (4, 3) and v.dtype == np.float32 and props[:3] == ["x", "y", "z"]
    assert f.tolist() == [[0, 1, 2], [0, 1, 3], [0, 2, 3], [0, 3, 1]]        # the quad is fan-triangulated
    out = tmp_path / "o.ply"
    st.write_ply_ascii(str(out), v + 1.0, f)
    v2, f2, _ = st.read_ply_ascii(str(out))
    assert np.allclose(v2, v + 1.0) and np.array_equal(f2, f)


def test_uniform_surface_sampling_is_area_weighted():
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [10, 0, 0], [10, 3, 0], [13, 0, 0]], np.float32)
    f = np.array([[0, 1, 2], [3, 4, 5]])            # areas 0.5 and 4.5
    pts = st.sample_points_uniformly(v, f, 20000, np.random.default_rng(1))
    assert pts.shape == (20000, 3) and pts.dtype == np.float32
    frac_big = float((pts[:, 0] > 5).mean())
    assert abs(frac_big - 0.9) < 0.01
    small = pts[pts[:, 0] < 5]
    assert (small[:, 0] >= 0).all() and (small[:, 1] >= 0).all() and (small[:, 0] + small[:, 1] <= 1 + 1e-6).all()
    assert abs(small[:, 0].mean() - 1 / 3) < 0.02   # uniform inside the triangle: centroid
