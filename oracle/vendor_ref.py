"""TEST INFRASTRUCTURE -- recipe that makes the UNMODIFIED reference package available as `oracle/_ref/`.

The reference (ace-tn v0.1.3) is pure Python: there is nothing to compile, "building" it means making it importable
where the checks run.  `/root/reference` does not exist on the GPU box, `oracle/_ref/` (git-ignored, NOT gpurun-ignored)
travels with the snapshot like the built `.so`.  Copied verbatim, never edited, never committed:

    /root/reference/acetn                                  -> oracle/_ref/acetn            (the package)
    /root/reference/tests/integration/{input,ipeps_gs}     -> oracle/_ref/tests/integration/...  (its own pins:
                                                              two converged states + energies.csv + their toml configs)

Consumers (checker side only -- SURVEY.md 8c): `tests/test_gpu_dropin.py` (the reference's Ipeps driven with
backend="b200" against the reference's own known answers and against its torch path), `tests/test_dropin_cpu.py`, and
`bench.py --impl reference` / `gpu_torch_baseline` (the reference's own DirectionalMover timed on the host cores / on the
same GPU).  The product (`acetn_b200/`) never imports it.

    python oracle/vendor_ref.py            # copy (idempotent)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
DEFAULT_SRC = os.environ.get("ACETN_REFERENCE", "/root/reference")

_PARTS = [("acetn", "acetn"),
          (os.path.join("tests", "integration", "input"), os.path.join("tests", "integration", "input")),
          (os.path.join("tests", "integration", "ipeps_gs"), os.path.join("tests", "integration", "ipeps_gs"))]


def available():
    """True when the vendored reference can be imported from oracle/_ref."""
    return os.path.isfile(os.path.join(DEST, "acetn", "__init__.py"))


def vendor(src=DEFAULT_SRC, force=False):
    """Copy the reference into oracle/_ref (no-op when the source tree is absent, e.g. on the GPU box)."""
    if not os.path.isdir(os.path.join(src, "acetn")):
        return available()
    for rel_src, rel_dst in _PARTS:
        s, d = os.path.join(src, rel_src), os.path.join(DEST, rel_dst)
        if os.path.isdir(d):
            if not force:
                continue
            shutil.rmtree(d)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so"))
    return available()


def import_path():
    """sys.path entry under which `import acetn` resolves to the vendored reference (or the original tree as a fallback
    inside the build container); None when neither exists."""
    if available():
        return DEST
    if os.path.isdir(os.path.join(DEFAULT_SRC, "acetn")):
        return DEFAULT_SRC
    return None


def enable():
    """Put the reference on sys.path (front).  Returns the path used, or None."""
    p = import_path()
    if p is not None and p not in sys.path:
        sys.path.insert(0, p)
    return p


if __name__ == "__main__":
    ok = vendor(force="--force" in sys.argv)
    print(f"oracle/_ref: {'ready' if ok else 'reference not available'} ({DEST})")
