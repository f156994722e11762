"""Recipe that makes the reference travel: copy exactly the /root/reference files the CLIP hot path imports into oracle/_ref/.

    python oracle/build_ref.py          (build container only; /root/reference is read-only and absent on the GPU box)

The set of files is DISCOVERED, not listed by hand: a child process imports the reference's CLIPApp / AdamW / CLIPEvaluator through
oracle/ref_loader.py, instantiates the three model branches' modules, and reports every module whose file lives under
/root/reference; those files (plus the package __init__.py chain) are copied verbatim, and their sha256 go to
oracle/_ref/MANIFEST.json.  oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so `bench.py --impl reference` and the `gpu_baseline` leg run the UNMODIFIED reference on the GPU box.
Called by __graft_entry__.build() when /root/reference is present."""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")

_PROBE = r"""
import json, os, sys
sys.path.insert(0, %r)
from oracle.ref_loader import import_reference
r = import_reference(%r)
import easynlp.modelzoo.models.clip.modeling_chineseclip, easynlp.modelzoo.models.clip.modeling_openclip, easynlp.modelzoo.models.clip.modeling_clip
import easynlp.modelzoo.models.roberta.modeling_roberta, easynlp.modelzoo.models.bert.modeling_bert
import easynlp.core.trainer, easynlp.core.optimizers, easynlp.core.evaluator, easynlp.core.predictor
files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(%r + os.sep)})
print("FILES=" + json.dumps(files))
"""


def build(verbose=False):
    if not os.path.isdir(os.path.join(REF, "easynlp")):
        return None
    env = dict(os.environ, HOME=os.environ.get("HOME", "/root"))
    p = subprocess.run([sys.executable, "-c", _PROBE % (ROOT, REF, REF)], capture_output=True, text=True, env=env)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("FILES=")]
    if p.returncode != 0 or not line:
        raise RuntimeError("reference import probe failed:\n" + p.stdout[-2000:] + p.stderr[-4000:])
    files = json.loads(line[0][6:])
    # package __init__ chain of every copied file (the two pre-seeded namespaces are deliberately NOT given their __init__)
    skip_init = {os.path.join(REF, "easynlp", "appzoo", "__init__.py"), os.path.join(REF, "easynlp", "appzoo", "clip", "__init__.py")}
    want = set(files)
    for f in files:
        d = os.path.dirname(f)
        while d.startswith(os.path.join(REF, "easynlp")):
            init = os.path.join(d, "__init__.py")
            if os.path.exists(init) and init not in skip_init:
                want.add(init)
            d = os.path.dirname(d)
    want -= skip_init
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    manifest = {}
    for f in sorted(want):
        rel = os.path.relpath(f, REF)
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(f, dst)
        manifest[rel] = hashlib.sha256(open(f, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "files": manifest}, fh, indent=1)
    # the copy must import on its own (this is what the GPU box will do)
    chk = subprocess.run([sys.executable, "-c",
                          "import sys; sys.path.insert(0, %r); from oracle.ref_loader import import_reference; r = import_reference(%r); print(r['CLIPApp'].__module__)" % (ROOT, OUT)],
                         capture_output=True, text=True, env=env)
    if chk.returncode != 0:
        raise RuntimeError("oracle/_ref does not import stand-alone:\n" + chk.stderr[-4000:])
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files, {sum(os.path.getsize(os.path.join(OUT, r)) for r in manifest) / 1e6:.2f} MB")
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
