"""TEST / BASELINE INFRASTRUCTURE - not product code.

Stages the UNMODIFIED reference files of the hot path under ``oracle/_ref/`` (git-ignored build output, travels to the GPU
box with the snapshot) so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the reference's own
``train_one_epoch`` on the reference's own module instead of the oracle restatement:

    classification/resnet/models/networks.py   (resnet50, ResNet, Bottleneck)
    classification/resnet/utils.py             (train_one_epoch)
    classification/vision_transformer/vit_model.py, utils.py

Run by ``__graft_entry__.build()`` when ``/root/reference`` exists (the build container); a no-op elsewhere.  Nothing here
is imported by ``deeplearning_b200``; no reference source enters the git history.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = [
    "classification/resnet/models/networks.py",
    "classification/resnet/utils.py",
    "classification/vision_transformer/vit_model.py",
    "classification/vision_transformer/utils.py",
]


def stage():
    """Copy the files (if the reference checkout is present). Returns True when oracle/_ref is usable."""
    if os.path.isdir(REF_ROOT):
        for rel in FILES:
            src, dst = os.path.join(REF_ROOT, rel), os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return available()


def available():
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


def _shim_matplotlib():
    """classification/*/utils.py import matplotlib.pyplot for a plotting helper the training loop never calls."""
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mp = types.ModuleType("matplotlib")
            mp.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mp
            sys.modules["matplotlib.pyplot"] = mp.pyplot


def load(project, module):
    """Import ``oracle/_ref/classification/<project>/<module>.py`` by path (e.g. load('resnet', 'models/networks'))."""
    import importlib.util

    _shim_matplotlib()
    path = os.path.join(DST, "classification", project, module + ".py")
    name = f"_ref_{project}_{module.replace('/', '_')}"
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("oracle/_ref staged" if stage() else "reference checkout not present: oracle/_ref not staged")
