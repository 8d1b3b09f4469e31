"""TEST / BENCH INFRASTRUCTURE -- recipe that installs the UNMODIFIED reference into oracle/_ref/.

    python oracle/build_ref.py [--force]

oracle/_ref/ is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the
installed package travels to the GPU box, where /root/reference does not exist.  What lands there:
  * the `keymorph` package, installed by pip from a scratch copy of /root/reference
    (`pip install --no-index --no-build-isolation --no-deps --target oracle/_ref`; the scratch copy is
    needed because the build writes egg-info into the source tree and /root/reference is read-only;
    --no-deps because ogb / outdated / torchio are not in the wheelhouse and not used on the path);
  * `scripts/` (register.py, pairwise_register_eval.py, script_utils.py ...): setup.py excludes it from the
    wheel; the drop-in test runs the stock `run_eval` from it against keymorph_b200;
  * `example_data_half/` (17 MB, the pair BASELINE config 1 names) for the config-1 parity test.
Called by __graft_entry__.build() when /root/reference is present; a no-op when oracle/_ref is up to date.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("KEYMORPH_REFERENCE", "/root/reference")
STAMP = os.path.join(DEST, ".installed_from")


def _src_signature():
    newest = 0.0
    for base, _, files in os.walk(os.path.join(SRC, "keymorph")):
        for f in files:
            newest = max(newest, os.path.getmtime(os.path.join(base, f)))
    return f"{SRC} {newest:.0f}"


def build(force=False, verbose=True):
    if not os.path.isdir(os.path.join(SRC, "keymorph")):
        if verbose:
            print(f"oracle/build_ref: {SRC} not present (GPU box?) -- keeping oracle/_ref as shipped")
        return os.path.isdir(os.path.join(DEST, "keymorph"))
    sig = _src_signature()
    if not force and os.path.exists(STAMP) and open(STAMP).read() == sig \
            and os.path.isdir(os.path.join(DEST, "keymorph")) and os.path.isdir(os.path.join(DEST, "scripts")):
        return True
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "src")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns("example_data*", "notebooks", ".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        how = "pip install --target"
        if r.returncode != 0 or not os.path.isdir(os.path.join(DEST, "keymorph")):
            # pip unavailable / build backend missing: the package is pure Python, a tree copy is the
            # same thing the wheel would have unpacked
            how = "tree copy (pip failed: %s)" % (r.stderr.strip().splitlines()[-1:] or ["?"])[0]
            shutil.copytree(os.path.join(SRC, "keymorph"), os.path.join(DEST, "keymorph"))
    shutil.copytree(os.path.join(SRC, "scripts"), os.path.join(DEST, "scripts"))
    init = os.path.join(DEST, "scripts", "__init__.py")
    if not os.path.exists(init):
        open(init, "w").close()
    if os.path.isdir(os.path.join(SRC, "example_data_half")):
        shutil.copytree(os.path.join(SRC, "example_data_half"), os.path.join(DEST, "example_data_half"))
    for base, dirs, files in os.walk(DEST):      # the source tree is read-only; the copy must be removable
        for n in dirs + files:
            os.chmod(os.path.join(base, n), 0o755 if n in dirs else 0o644)
    with open(STAMP, "w") as f:
        f.write(sig)
    if verbose:
        print(f"oracle/build_ref: reference installed into {DEST} ({how})")
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
