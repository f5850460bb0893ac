"""TEST INFRASTRUCTURE — install the UNMODIFIED reference next to the oracle (``oracle/_ref``).

    python -m oracle.build_ref

The reference is a pure-Python package: "building" it is packing its importable modules from
where they lie (``/root/reference``, build container only) into one importable archive,
``oracle/_ref/reference.zip`` (package ``frldistml.scaffold``, the name its own tests use).
``oracle/_ref`` is git-ignored (no reference source enters the history) but travels to the GPU
box with the snapshot like a built ``.so``, so that there ``bench.py --impl reference`` and the
``cpu_baseline`` leg time the reference's own ``SolverWorker._pass_one_minibatch`` on the host
cores (``kind: "reference"``) instead of the restatement in ``oracle/ref_loop.py``
(``kind: "port"``, the fallback when ``oracle/_ref`` is absent).  Nothing in the product imports it.
"""
import os
import sys
import zipfile

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "reference.zip")
SKIP = {"tests", ".git", "__pycache__", ".github"}


def build(verbose: bool = True) -> str:
    """Pack the reference's modules; returns the archive path ('' if there is neither a source
    tree nor an archive from an earlier build)."""
    if not os.path.isdir(SRC):
        return DST if os.path.exists(DST) else ""
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    n = 0
    with zipfile.ZipFile(DST + ".tmp", "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("frldistml/__init__.py", "")
        for root, dirs, files in os.walk(SRC):
            dirs[:] = [d for d in dirs if d not in SKIP]
            for f in files:
                if f.endswith(".py"):
                    full = os.path.join(root, f)
                    z.write(full, os.path.join("frldistml", "scaffold", os.path.relpath(full, SRC)))
                    n += 1
    os.replace(DST + ".tmp", DST)
    if verbose:
        print("packed the unmodified reference (%d modules) into %s" % (n, DST))
    return DST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
