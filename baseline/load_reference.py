"""Locates and imports the UNMODIFIED reference (ericbrec/BSpy 5.0.1) for bench.py's CPU arm.

Search order: ``baseline/_ref`` (``pip install --no-deps --target baseline/_ref`` of the reference, done by
``__graft_entry__.build()`` in the build container; git-ignored, travels to the GPU box with the snapshot), then
``/root/reference`` (only exists in the build container).  The reference's ``bspy/__init__.py:28-29`` hard-imports its
tkinter / PyOpenGL viewer, which no headless box has, so inert stand-ins for those GUI modules are registered before
the import; nothing of the evaluation path touches them.  Nothing here is used by the product (``bspy_b200``): only
``bench.py --impl reference`` and the ``cpu_baseline`` leg call ``load()``.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


def install(source="/root/reference", target=os.path.join(HERE, "_ref")):
    """pip-install the reference tree into baseline/_ref (offline, --no-deps: numpy / scipy are in the image, the GUI
    dependencies are not needed).  The source tree is read-only, so the build runs from a copy under /tmp."""
    import shutil
    import subprocess
    import tempfile
    if not os.path.isdir(source):
        return False, f"{source} does not exist"
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "src")
        shutil.copytree(source, work, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade",
               "--find-links", "/opt/wheelhouse", "--target", target, work]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    ok = p.returncode == 0 and os.path.isdir(os.path.join(target, "bspy"))
    return ok, (p.stdout or "").strip()[-400:]


def _stub_gui_modules():
    class _Inert:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Inert()

        def __call__(self, *a, **k):
            return _Inert()

    for name in ("tkinter", "tkinter.ttk", "tkinter.colorchooser", "tkinter.filedialog",
                 "OpenGL", "OpenGL.GL", "OpenGL.GLU", "OpenGL.GL.shaders", "pyopengltk"):
        if name in sys.modules:
            continue
        m = types.ModuleType(name)
        m.__all__ = []
        m.__path__ = []
        m.__getattr__ = lambda attr, _n=name: type(attr, (_Inert,), {})
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)


def load():
    """Returns ``(bspy_module, root)`` of the unmodified reference, or ``(None, reason)``."""
    reasons = []
    for root in CANDIDATES:
        if not os.path.isdir(os.path.join(root, "bspy")):
            reasons.append(f"{root}: no bspy package")
            continue
        _stub_gui_modules()
        sys.path.insert(0, root)
        try:
            sys.modules.pop("bspy", None)
            import bspy
            if not os.path.abspath(bspy.__file__).startswith(os.path.abspath(root)):
                raise ImportError(f"imported {bspy.__file__}, not the reference under {root}")
            if not hasattr(bspy, "Spline") or not hasattr(bspy.Spline, "jacobian"):
                raise ImportError("package has no Spline.jacobian")
            return bspy, root
        except Exception as exc:  # pragma: no cover - depends on the box
            reasons.append(f"{root}: {type(exc).__name__}: {exc}")
            sys.path.remove(root)
            for k in [k for k in sys.modules if k == "bspy" or k.startswith("bspy.")]:
                sys.modules.pop(k, None)
    return None, "; ".join(reasons)
