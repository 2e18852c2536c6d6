"""Node-feature fixtures of the bundled MUSAE datasets (test infrastructure; run in the build container).

    python oracle/make_golden_features.py            # -> tests/golden/{twitch,fb}_features.npz

The GPU box has no /root/reference, so the real binary bag-of-features matrices of twitch-DE
(9,498 x 2,514 non-zero columns) and musae-facebook (22,470 x 4,714) — public dataset files shipped next to
the reference's loaders — travel as compressed CSR index lists.  Processing restates
/root/reference/twitch/data.py:73-84 (fb/data.py: same code): features[node, feats] = 1 for node < n over a
zero matrix of the dataset's nominal width, then all-zero columns are dropped.  The dense matrix is rebuilt by
``edge_proposal_sets_b200.data`` when the CSV/JSON files are not in the working directory.
"""
import json
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SPEC = {"twitch": ("twitch/musae_DE_features.json", 9498, 3170),
        "fb": ("fb/musae_facebook_features.json", 22470, 4714)}


def main():
    for name, (path, n, width) in SPEC.items():
        with open(os.path.join(REF, path)) as f:
            j = json.load(f)
        feats = np.zeros((n, width), dtype=np.uint8)
        for node, fl in j.items():
            if int(node) < n:
                feats[int(node), np.asarray(fl, dtype=int)] = 1
        x = feats[:, feats.sum(0) != 0]
        rows, cols = np.nonzero(x)
        indptr = np.zeros(n + 1, dtype=np.int32)
        np.add.at(indptr, rows + 1, 1)
        np.savez_compressed(os.path.join(OUT, f"{name}_features.npz"), n=n, width=x.shape[1],
                            indptr=np.cumsum(indptr).astype(np.int32), indices=cols.astype(np.uint16),
                            checksum=np.int64(int((rows.astype(np.int64) * 31 + cols).sum())))
        print(name, x.shape, int(x.sum()), "non-zeros")


if __name__ == "__main__":
    main()
