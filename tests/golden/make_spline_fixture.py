"""Extracts the T106A blade coordinate table the reference's spline known-answer test integrates over
(`src/core/spline.zig:306-514`, "T106 blade coordinate integration": coordinates over chord from R. D. Stieger's thesis,
Table I-2; chord 198 mm; suction + pressure surface length 264.7 mm + 230.0 mm, tolerance 1e-2) into
``tests/golden/t106a_blade_table.npz``.  Run in the build container (needs /root/reference); the GPU box only reads the
committed fixture."""
import os
import re

import numpy as np

REF = "/root/reference/src/core/spline.zig"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    text = open(REF).read()
    body = text[text.index('test "T106 blade coordinate integration"'):]
    body = body[:body.index("const chord")]
    pts = np.array([[float(a), float(b)] for a, b in re.findall(r"\.\{\s*(-?[0-9.]+),\s*(-?[0-9.]+)\s*\}", body)])
    assert len(pts) > 100 and np.array_equal(pts[0], pts[-1])  # a closed contour
    np.savez_compressed(os.path.join(HERE, "t106a_blade_table.npz"), points_over_chord=pts, chord=0.198, surface_length=0.2647 + 0.2300, tolerance=1e-2)
    print(len(pts), "points")
