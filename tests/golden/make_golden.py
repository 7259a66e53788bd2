"""Regenerates tests/golden/obj_1.npz from the reference's golden object test/obj_1.mat.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The fixture is DATA copied from the reference's own saved run (fp64, 35x35 grid, 100 controls,
N = 130; constructor settings in test/obj_1.txt) — the only stored output the reference ships.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.mcos import load_mcos_properties  # noqa: E402


def main(ref="/root/reference/test/obj_1.mat"):
    p = load_mcos_properties(ref)
    u_star = np.asarray(p["u_star"], dtype=np.float64)
    J_star = np.asarray(p["J_star"], dtype=np.float64)
    u_vals, u_code = np.unique(u_star, return_inverse=True)
    assert len(u_vals) < 256
    u_code = u_code.reshape(u_star.shape).astype(np.uint8)
    assert np.array_equal(u_vals[u_code], u_star)
    sha_u = hashlib.sha256(u_star.tobytes(order="F")).hexdigest()
    sha_J = hashlib.sha256(J_star.tobytes(order="F")).hexdigest()
    out = os.path.join(HERE, "obj_1.npz")
    np.savez_compressed(
        out,
        A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=p["N"], dx=p["dx"], du=p["du"],
        x_min=p["x_min"], x_max=p["x_max"], u_min=p["u_min"], u_max=p["u_max"],
        X1_mesh=p["X1_mesh"], X2_mesh=p["X2_mesh"],
        u_vals=u_vals, u_code=u_code, J_star=J_star,
        sha256_u_star=np.array(sha_u), sha256_J_star=np.array(sha_J),
    )
    print(out, os.path.getsize(out), "bytes; sha u_star", sha_u[:16], "J_star", sha_J[:16])


if __name__ == "__main__":
    main(*sys.argv[1:])
