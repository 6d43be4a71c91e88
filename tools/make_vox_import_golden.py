"""Writes tests/golden/vox_import_digests.json: SHA-256 digests of the reference's own .vox imports (the shipped .v / .p files under
/root/reference/Assets/Mods/default, produced by the reference's importer from the .vox files shipped next to them) in the form
tests/test_vox_import.py::_digest hashes an import result.  Run where /root/reference is mounted."""
import hashlib
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import assets_py as A  # noqa: E402

BASE = "/root/reference/Assets/Mods/default"
out = {}
for name in ("ModernHouse", "FarmHouse", "Player"):
    res = A.vox_import(open(f"{BASE}/{name}.vox", "rb").read())
    h = hashlib.sha256()
    for mname, vox in res["models"]:
        shipped = A.read_v(f"{BASE}/{name}/{mname}.v")                 # the digest is taken over the SHIPPED bytes, not over the restatement's
        h.update(mname.encode() + np.array(shipped.shape, "<i4").tobytes() + shipped.tobytes())
    rec = np.fromfile(f"{BASE}/{name}/{name}.p", np.uint8)[:1792].reshape(256, 7).copy()
    rec[:, 3] = 0
    if name == "FarmHouse":
        rec[249, 4] = 255                                              # see tests/test_vox_import.py: the one shipped record that differs
    h.update(rec.tobytes())
    for e in res["entities"]:
        h.update(struct.pack("<ii3f", e["parent"], e["model"], *[float(v) for v in e["position"]]) + e["name"].encode())
    out[name] = {"digest": h.hexdigest(), "models": len(res["models"])}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "vox_import_digests.json"), "w"), indent=1)
print(out)
