"""TEST INFRASTRUCTURE — writes tests/golden/indexed/* with the UNMODIFIED reference's
IndexedDatasetWriter (reference storage_layers/dataset.py:607-655), so the committed .idx/.bin
pairs are what the reference itself produces.  Run in the build container only:

    python oracle/make_indexed_golden.py
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "indexed")

# name -> (dtype, frame shape, number of frames, seed)
FILES = {
    "mono": ("uint8", (3, 8, 8), 11, 0),
    "pose": ("float32", (6,), 11, 1),
    "label": ("int64", (1,), 11, 2),
    "wide": ("float64", (2, 3, 5), 4, 3),
    "scalar": ("int16", (), 5, 4),
}


def frames_of(name):
    dtype, shape, n, seed = FILES[name]
    rs = np.random.RandomState(seed)
    return [np.asarray(rs.randn(*shape) * 40).astype(dtype) for _ in range(n)]


def main():
    import_reference()
    from frldistml.scaffold.indexed_dataset import IndexedDatasetWriterFactory
    from frldistml.scaffold.storage import StoragePath
    from frldistml.scaffold.storage_layers.posix_storage import PosixIndexedDatasetReader
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name in FILES:
        idx, binf = os.path.join(OUT, name + ".idx"), os.path.join(OUT, name + ".bin")
        with IndexedDatasetWriterFactory.get(idxfile=StoragePath(idx), binfile=StoragePath(binf)) as w:
            for f in frames_of(name):
                w.push_back(f)
        reader = PosixIndexedDatasetReader(idxfile=StoragePath(idx), binfile=StoragePath(binf))
        # what the reference's own reader returns for these files: recorded for the GPU box,
        # where the reference is absent
        manifest[name] = {"len": len(reader), "dtype": str(reader.dtype), "framesize": int(reader.framesize),
                          "size": [int(d) for d in reader.size],
                          "frame_sums": [float(np.asarray(reader[i], dtype=np.float64).sum())
                                         for i in range(len(reader))],
                          "first_frame": np.asarray(reader[0]).ravel().tolist()}
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
