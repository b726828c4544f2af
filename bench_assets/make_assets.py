"""Generates the reference-cooked assets the product's facade scenes load (run in the container that has /root/reference built
into oracle/_ref): convex hull cooking (ConvexHullBuilder) and mesh cooking (AABBTreeBuilder + codecs) are host-side and out of
scope, so the reference's output is used verbatim.

  convex_vs_mesh.b2js : [terrain MeshShape of PerformanceTest/ConvexVsMeshScene.h, its 5 point ConvexHullShape]
  pile_hulls.b2js     : the palette of 256 random 12 point hulls of the Pile scene
"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
sys.path.insert(0, os.path.join(HERE, ".."))
import refharness as R  # noqa: E402

L = R.ref_lib("det")
L.jref_dump_cooked_shapes.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
w = R.RefWorld("convex_vs_mesh", 1)
print("convex_vs_mesh:", L.jref_dump_cooked_shapes(w.h, os.path.join(HERE, "convex_vs_mesh.b2js").encode()))
w = R.RefWorld("pile", 256, 8)
print("pile_hulls:", L.jref_dump_cooked_shapes(w.h, os.path.join(HERE, "pile_hulls.b2js").encode()))
