#!/usr/bin/env python
"""Regenerates the CPU golden fixtures from the REFERENCE's own host code.

Needs oracle/_ref/libref_host.so (built by `make -C oracle` from /root/reference sources, so this
script only runs where the reference tree is mounted).  Output: tests/golden/structure.npz and
tests/golden/structure_meta.json, both committed.

  structure.npz   for n in (1*, 2, 3, 4, 5, 7, 16): the reference generator's file -> the
                  reference reader's Entry[] -> the reference build_csr_struct arrays
                  (* n = 1 is skipped if the reference generator cannot represent it)
  meta.json       sha256 of the generated .mtx files and of the bundled matrix/example81x81.mtx,
                  plus its CSR checksum vectors (sum of row_ptr / col / values)
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import orc  # noqa: E402

REF_ROOT = os.environ.get("REF_ROOT", "/root/reference")


def main():
    R = orc.RefHost()
    td = tempfile.mkdtemp()
    arrays, meta = {}, {"files": {}, "generator": "reference write_matrix_market_stencil5 / load_matrix_market / build_csr_struct"}
    for n in (2, 3, 4, 5, 7, 16):
        path = os.path.join(td, "s%d.mtx" % n)
        R.write_stencil(n, path)
        meta["files"]["stencil_%d" % n] = hashlib.sha256(open(path, "rb").read()).hexdigest()
        m, ent = R.load(path)
        rp, ci, va = R.build_csr(m)
        arrays["n%d_entries_row" % n] = ent["row"].copy()
        arrays["n%d_entries_col" % n] = ent["col"].copy()
        arrays["n%d_entries_val" % n] = ent["value"].copy()
        arrays["n%d_row_ptr" % n] = rp
        arrays["n%d_col" % n] = ci
        arrays["n%d_val" % n] = va
        meta["n%d" % n] = {"rows": m.rows, "cols": m.cols, "nnz": m.nnz, "grid_size": m.grid_size}
    bundled = os.path.join(REF_ROOT, "matrix", "example81x81.mtx")
    raw = open(bundled, "rb").read()
    m, ent = R.load(bundled)
    rp, ci, va = R.build_csr(m)
    meta["bundled81"] = {
        "sha256": hashlib.sha256(raw).hexdigest(), "rows": m.rows, "cols": m.cols, "nnz": m.nnz,
        "grid_size": m.grid_size, "row_ptr_sum": int(rp.astype(np.int64).sum()),
        "col_sum": int(ci.astype(np.int64).sum()), "val_sum": float(va.sum()),
        "csr_sha256": hashlib.sha256(rp.tobytes() + ci.tobytes() + va.tobytes()).hexdigest(),
        "entries_sha256": hashlib.sha256(ent.tobytes()).hexdigest(),
    }
    np.savez_compressed(os.path.join(HERE, "structure.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "structure_meta.json"), "w"), indent=1, sort_keys=True)
    print("wrote structure.npz (%d arrays) and structure_meta.json" % len(arrays))


if __name__ == "__main__":
    main()
