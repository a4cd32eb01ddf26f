"""Structural fingerprints of the dolfin-written HDF5 files in the reference's test data, and the text of the XDMF the
reference itself writes for a ``write_checkpoint`` series (build container only).

    python tests/golden/make_h5_structure_golden.py        ->  tests/golden/h5_structure.json

``tests/h5struct.py`` walks the on-disk structures of ``/root/reference/tests/test_data/**/*.h5`` (legacy dolfin,
HDF5 1.12, ``libver=earliest``) and keeps what is format, not content; ``create_checkpoint_xdmf_file``
(``postprocessing_h5py_common.py:594-682``) is imported unmodified from ``/root/reference/src`` and run once.
``tests/test_h5_structure.py`` holds ``vasp_b200.h5lite.H5Writer`` / ``io_dolfin.CheckpointWriter`` to both.
"""
import json
import sys
import tempfile
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from tests import h5struct as hs  # noqa: E402
from tests.golden.make_reference_goldens import install_shims  # noqa: E402

REF_DATA = Path("/root/reference/tests/test_data")


def main() -> None:
    out = {"files": {}}
    for p in sorted(REF_DATA.rglob("*.h5")):
        fp = hs.fingerprint(p)
        out["files"][str(p.relative_to(REF_DATA))] = {
            "superblock": fp["superblock"],
            "objects": {k: hs.object_signature(v) for k, v in fp["objects"].items()}}
    install_shims()
    import importlib
    common = importlib.import_module("vasp.postprocessing.postprocessing_h5py.postprocessing_h5py_common")
    with tempfile.TemporaryDirectory() as d:
        for att in ("Scalar", "Vector"):
            common.create_checkpoint_xdmf_file(3, 0.25, 0.5, 8 * 11, 17, att, f"Q{att}", Path(d))
            out[f"checkpoint_xdmf_{att}"] = (Path(d) / f"Q{att}.xdmf").read_text()
    (HERE / "h5_structure.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print(f"wrote {HERE / 'h5_structure.json'}: {len(out['files'])} files")


if __name__ == "__main__":
    main()
