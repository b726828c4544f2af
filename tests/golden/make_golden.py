"""Generates the golden fixtures of tests/golden/*.npz from the REFERENCE (oracle/_ref/libjoltref_det.so = the unmodified
jrouwe/JoltPhysics sources compiled with CROSS_PLATFORM_DETERMINISTIC, see oracle/Makefile). Run in the build container:

    python tests/golden/make_golden.py

Each fixture holds, for one benchmark scene built by the reference, the body state (centre of mass position, rotation, linear and
angular velocity, active flag) after k Update(1/60, 1) calls for several k, plus the candidate pair set and the contact cache
summary (manifolds per pair, points per manifold) at those steps. The -m gpu tests build the SAME scene through the facade (no
reference needed on the GPU box) and must reproduce these states; test_golden.py also pins the fixtures to the oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness as R  # noqa: E402

CASES = [("pyramid", 6, 0, (1, 30, 120)), ("convex_vs_mesh", 2, 0, (1, 60, 150)), ("pile", 600, 15, (1, 60, 150)), ("max_bodies", 300, 0, (1, 20)),
         ("feature", 8, 0, (1, 30, 90, 200)),  # feature 8 = the zoo: every motion type / body flag / override in one world
         ("feature", 9, 0, (1, 40, 100, 220)),  # feature 9 = ScaledShape / RotatedTranslatedShape around every convex leaf type
         ("feature", 10, 0, (1, 40, 100, 220)),  # feature 10 = CylinderShape plain / scaled / rotated against the other convex shapes
         ("feature", 11, 0, (1, 40, 120, 250))]  # feature 11 = PointConstraint / DistanceConstraint / HingeConstraint (chain, rope, cloth = one large island, hinges)


def snapshot(world):
    """State + sorted candidate pair rows + sorted (body1, body2, sub1, sub2, num_points) rows of the contact cache."""
    s = world.state()
    p, n_p, m, n_m = world.cache()
    return {"pos": s.pos, "rot": s.rot, "lin": s.lin, "ang": s.ang, "active": (s.active_index != 0xffffffff),
            "pairs": R.sorted_rows(world.find_pairs()), "cache": R.cache_rows(R.cache_summary(p, n_p, m, n_m))}


def main():
    for scene, p0, p1, steps in CASES:
        ref = R.RefWorld(scene, p0, p1)
        out = {}
        done = 0
        for k in steps:
            while done < k:
                ref.step()
                done += 1
            for name, v in snapshot(ref).items():
                out[f"s{k}_{name}"] = v
        out["steps"] = np.array(steps)
        path = os.path.join(HERE, f"{scene}_{p0}_{p1}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes")
        ref.close()


if __name__ == "__main__":
    main()
