"""Where the car / track / cfg data of a reference checkout lives (the `basePath` of PyProjectD.createSimulator).

The product reads the reference's own on-disk formats (.ini / .lut / .rto / surfaces.bin / spline.cache / collider.bin);
this module only answers "which directory" when the caller did not say."""
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def default_base():
    """PD_BASE_PATH, else the in-tree content copy (assets/_base, made by __graft_entry__.build()), else /root/reference."""
    for p in (os.environ.get("PD_BASE_PATH"), os.path.join(_ROOT, "assets", "_base"), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "content", "cars")) and os.path.isdir(os.path.join(p, "cfg")):
            return p
    raise FileNotFoundError("no base directory with cfg/ and content/ found (set PD_BASE_PATH or run __graft_entry__.build())")
