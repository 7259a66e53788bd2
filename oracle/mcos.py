"""Decoder for the reference's only golden vector, ``test/obj_1.mat``.  TEST INFRASTRUCTURE.

``obj_1.mat`` is a MATLAB 5.0 MAT-file holding a saved ``Dynamic_Solver`` *object* (MCOS).
scipy cannot open MCOS objects directly, but the property values sit in the
``__function_workspace__`` blob, which is itself a MAT5 stream (recipe: SURVEY.md 4.3).

Used by ``tests/golden/make_golden.py`` (run in the build container, where ``/root/reference``
exists) to produce the committed fixture ``tests/golden/obj_1.npz``.
"""
import io
import struct

import numpy as np


def load_mcos_properties(path):
    """Return {property name: ndarray} of the single object stored in ``path``."""
    import scipy.io
    from scipy.io.matlab._mio5 import MatFile5Reader

    raw = open(path, "rb").read()
    d = scipy.io.loadmat(path)
    fw = d["__function_workspace__"].tobytes()
    stream = io.BytesIO(raw[:128] + fw[8:])
    rd = MatFile5Reader(stream)
    rd.initialize_read()
    stream.seek(128)
    hdr, _ = rd.read_var_header()
    arr = rd.read_var_array(hdr, process=False)
    cells = arr["MCOS"][0, 0]["_ObjectMetadata"][0]
    meta = cells[0, 0].tobytes()
    _ver, nstr = struct.unpack_from("<II", meta, 0)
    offs = struct.unpack_from("<8I", meta, 8)
    names = [s.decode() for s in meta[40:offs[0]].split(b"\x00")[:nstr]]
    seg = np.frombuffer(meta[offs[3]:offs[4]], dtype="<u4")
    nprop = int(seg[2])
    props = {}
    for k in range(nprop):
        name_idx, _kind, val_idx = (int(x) for x in seg[3 + 3 * k:6 + 3 * k])
        props[names[name_idx - 1]] = np.asarray(cells[val_idx + 2, 0])
    return props
