"""Install the UNMODIFIED reference forward into baseline/_ref (git-ignored, shipped to the GPU box by gpurun).

The reference (mehrdad78/SUNet_TF) is a directory of Python scripts without setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to build.  The install is therefore a verbatim copy of
the three files the forward path consists of (model/SUNet.py, model/SUNet_detail.py, training.yaml) - nothing else of
the tree is needed by `SUNet_model(opt)(x)`.  baseline/_ref never enters git history (.gitignore) and nothing under
sunet_tf_b200/ reads it: it is used by `bench.py --impl reference`, bench.py's cpu_baseline leg and the tests, through
oracle/reference_loader.py, as the live reference (kind "reference" instead of the oracle "port").

  python baseline/install_ref.py [--src /root/reference]
"""
import argparse
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("model/SUNet.py", "model/SUNet_detail.py", "training.yaml")


def install(src="/root/reference", dest=DEST):
    """Returns the install directory, or None when the reference tree is not present (GPU box: uses what was shipped)."""
    if not os.path.isfile(os.path.join(src, "model", "SUNet_detail.py")):
        return dest if os.path.isfile(os.path.join(dest, "model", "SUNet_detail.py")) else None
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    print(install(ap.parse_args().src))
