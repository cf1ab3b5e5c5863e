#!/usr/bin/env python3
"""Extracts the known-answer colour-difference vectors of the reference's own unit tests
(/root/reference/test/tst_ColourDifference.h:26-35, 49-56, 73-108) into colour_vectors.json.

The numbers are data (Sharma et al. 2005 for CIEDE2000); no reference code is copied.
Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_colour_vectors.py
"""
import json
import os
import re
import sys

REF = os.environ.get("MOSAIC_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "test", "tst_ColourDifference.h")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "colour_vectors.json")

TESTS = {"RGBEuclidean": ("rgb_euclidean", 1e-8), "CIE76": ("cie76", 1e-8), "CIEDE2000": ("ciede2000", 1e-4)}
ROW = re.compile(r"\{\{([^}]*)\},\s*\{([^}]*)\},\s*([-0-9.eE]+)\}")


def main():
    text = open(SRC, encoding="utf-8", errors="replace").read()
    out = {"source": "test/tst_ColourDifference.h", "sets": {}}
    for name, (key, tol) in TESTS.items():
        m = re.search(r"TEST\(ColourDifference,\s*%s\)(.*?)for \(const auto" % name, text, re.S)
        if not m:
            sys.exit("test %s not found" % name)
        rows = []
        for a, b, d in ROW.findall(m.group(1)):
            rows.append({"first": [float(v) for v in a.split(",")],
                         "second": [float(v) for v in b.split(",")], "difference": float(d)})
        out["sets"][key] = {"tolerance": tol, "vectors": rows}
    json.dump(out, open(OUT, "w"), indent=1)
    print({k: len(v["vectors"]) for k, v in out["sets"].items()})


if __name__ == "__main__":
    main()
