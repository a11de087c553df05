"""Headless stand-ins for the GUI / convenience dependencies of the reference's CLI scripts, so that
eval_nolearned.py and shape_transfer.py run UNMODIFIED on a machine without open3d / easydict:

    import deformationpyramid_b200 as ndp
    ndp.headless.install()          # open3d (mesh I/O + surface sampling, no windows), easydict
    ndp.install_as_model()          # model.nets / model.loss / model.registration -> the sm_100a mirror
    runpy.run_path("shape_transfer.py", run_name="__main__")

open3d surface used by shape_transfer.py:4,69-83,166-168: io.read_triangle_mesh (ASCII PLY),
TriangleMesh.compute_vertex_normals / sample_points_uniformly / vertices, PointCloud.points /
paint_uniform_color, utility.Vector3dVector, visualization.draw_geometries (no-op).  model/geometry.py:4 and
utils/benchmark_utils.py:5 only import the module.  Real packages, when installed, are left alone.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from .config import AttrDict


class _PointCloud:
    def __init__(self, points):
        self.points = np.asarray(points, dtype=np.float64)
        self.colors = None

    def paint_uniform_color(self, color):
        self.colors = np.tile(np.asarray(color, dtype=np.float64), (len(self.points), 1))
        return self


class _TriangleMesh:
    def __init__(self, vertices, triangles):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.triangles = np.asarray(triangles, dtype=np.int64)
        self.vertex_normals = None

    def compute_vertex_normals(self):
        return self                                   # shading only (shape_transfer.py:70,79)

    def sample_points_uniformly(self, number_of_points=100, seed=None):
        from .shape_transfer import sample_points_uniformly
        rng = np.random.default_rng(_state["seed"] if seed is None else seed)
        _state["seed"] += 1
        return _PointCloud(sample_points_uniformly(np.asarray(self.vertices, np.float32), self.triangles,
                                                   int(number_of_points), rng))


_state = {"seed": 0, "drawn": 0}


def _read_triangle_mesh(path, *a, **k):
    from .shape_transfer import read_ply_ascii
    v, f, _ = read_ply_ascii(path)
    return _TriangleMesh(v, f)


def _draw_geometries(geoms, *a, **k):
    _state["drawn"] += 1                                # a window in the reference; nothing here


def open3d_module() -> types.ModuleType:
    o3d = types.ModuleType("open3d")
    o3d.__doc__ = "headless stand-in (deformationpyramid_b200.headless)"
    o3d.io = types.ModuleType("open3d.io")
    o3d.io.read_triangle_mesh = _read_triangle_mesh
    o3d.geometry = types.ModuleType("open3d.geometry")
    o3d.geometry.TriangleMesh = _TriangleMesh
    o3d.geometry.PointCloud = _PointCloud
    o3d.utility = types.ModuleType("open3d.utility")
    o3d.utility.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64)
    o3d.utility.Vector3iVector = lambda a: np.asarray(a, dtype=np.int64)
    o3d.visualization = types.ModuleType("open3d.visualization")
    o3d.visualization.draw_geometries = _draw_geometries
    return o3d


def install(force: bool = False, seed: int = 0) -> None:
    """Register the stand-ins for every module that is not importable (or for all with force=True)."""
    import importlib
    _state["seed"], _state["drawn"] = int(seed), 0

    def missing(name):
        if force:
            return True
        try:
            importlib.import_module(name)
            return False
        except ImportError:
            return True

    if missing("open3d"):
        o3d = open3d_module()
        for m in (o3d, o3d.io, o3d.geometry, o3d.utility, o3d.visualization):
            sys.modules[m.__name__] = m
    if missing("easydict"):
        ed = types.ModuleType("easydict")
        ed.EasyDict = AttrDict
        sys.modules["easydict"] = ed
