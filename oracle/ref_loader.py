"""Import the UNMODIFIED reference (alibaba/EasyNLP) hot-path modules.  TEST / BENCH INFRASTRUCTURE ONLY -- never imported by
easynlp_b200.

Where it comes from: /root/reference when that exists (the build container), else oracle/_ref/ -- a verbatim copy of exactly
the reference files this path imports, made by oracle/build_ref.py (git-ignored, travels to the GPU box with the snapshot like the
built .so files; sha256 of every file recorded in oracle/_ref/MANIFEST.json).  The reference is pure Python on this path, so
"building" it is a file copy.

`easynlp.appzoo.__init__` eagerly imports every application (and, through them, `imp`, `datasets.list_datasets`, ftfy, rouge, ...
which do not exist here), so bare namespace modules are pre-seeded for `easynlp.appzoo` and `easynlp.appzoo.clip`;
the reference files themselves are untouched."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIVE = "/root/reference"
REF_COPY = os.path.join(HERE, "_ref")


def reference_root(prefer_copy: bool = False):
    """path that holds the `easynlp/` package, or None"""
    cands = [REF_COPY, REF_LIVE] if prefer_copy else [REF_LIVE, REF_COPY]
    for c in cands:
        if os.path.isdir(os.path.join(c, "easynlp", "appzoo", "clip")):
            return c
    return None


def import_reference(root=None):
    """-> dict(CLIPApp, AdamW, CLIPEvaluator, root).  Raises ImportError when neither location holds the reference."""
    root = root or reference_root()
    if root is None:
        raise ImportError("reference not available: neither /root/reference nor oracle/_ref (run `python oracle/build_ref.py` in the build container)")
    os.environ.setdefault("HOME", "/root")
    if root not in sys.path:
        sys.path.insert(0, root)
    for name, sub in (("easynlp.appzoo", "easynlp/appzoo"), ("easynlp.appzoo.clip", "easynlp/appzoo/clip")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(root, sub)]
            sys.modules[name] = m
    from easynlp.appzoo.clip.model import CLIPApp
    from easynlp.core.optimizers import AdamW
    out = {"CLIPApp": CLIPApp, "AdamW": AdamW, "root": root}
    try:
        from easynlp.appzoo.clip.evaluator import CLIPEvaluator
        out["CLIPEvaluator"] = CLIPEvaluator
    except Exception:      # the evaluator drags in more of the package; not needed by the timing legs
        out["CLIPEvaluator"] = None
    return out
