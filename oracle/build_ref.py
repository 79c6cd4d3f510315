"""TEST INFRASTRUCTURE — recipe for oracle/_ref/: the reference's own pure-Python hot-path modules, copied UNMODIFIED
from /root/reference (where they lie in the build container) so that they travel to the GPU box with the repo snapshot.

oracle/_ref/ is build output: listed in .gitignore (never committed), not in .gpurunignore. The copy is made by
`__graft_entry__.build()` whenever /root/reference is present; on the GPU box only the prebuilt copy is used.
Consumers: `bench.py --impl reference` (times the reference's `sampling.iterate_state_probability`, sampling.py:29-60,
through oracle/ref_shim.py) and tests that pin the oracle against the reference. Nothing in cy2path_b200/ reads it.

Files (all pure Python; their third-party imports are numpy / scipy / torch / joblib / tqdm, present in the image; the
`scvelo` names utils.py imports are stubbed by oracle/ref_shim.py):
  cy2path/sampling.py  cy2path/simulation.py  cy2path/utils.py  cy2path/models/{__init__,_FHMM,_LSSM,_NSSM,methods,trainer}.py
"""
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CY2PATH_REFERENCE", "/root/reference")
REF_DST = os.path.join(_HERE, "_ref")
FILES = ["cy2path/sampling.py", "cy2path/simulation.py", "cy2path/utils.py", "cy2path/models/__init__.py",
         "cy2path/models/_FHMM.py", "cy2path/models/_LSSM.py", "cy2path/models/_NSSM.py", "cy2path/models/methods.py",
         "cy2path/models/trainer.py"]


def build():
    """Copy the modules when the reference tree is present; returns the destination (or None when there is neither a
    source tree nor a previous copy)."""
    if os.path.isdir(os.path.join(REF_SRC, "cy2path")):
        for rel in FILES:
            dst = os.path.join(REF_DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(REF_SRC, rel), dst)
            os.chmod(dst, 0o644)
        with open(os.path.join(REF_DST, "README"), "w") as f:
            f.write("Unmodified copies of aron0093/cy2path hot-path modules (build output of oracle/build_ref.py). "
                    "Not part of the repository history.\n")
    return REF_DST if available() else None


def available():
    return all(os.path.exists(os.path.join(REF_DST, rel)) for rel in FILES)


if __name__ == "__main__":
    print(build())
