#!/usr/bin/env python
"""Stage the UNMODIFIED reference package into the git-ignored oracle/_ref/.

    python -m oracle.stage_ref [REFERENCE_CHECKOUT]      (default /root/reference)

TEST / BENCH INFRASTRUCTURE.  The reference (nicoaira/GINFINITY 1.2.1) is pure
Python: nothing to compile, the "build" of oracle/_ref is a verbatim copy of

    src/ginfinity/            the package, bundled checkpoint included
    tests/*.py                its own test-suite (run under the C-ABI binding by
                              tests/test_gpu_reference.py)
    tests/rouskin_sample_6k.tsv   BASELINE configs[0] input (C1)
    LICENSE, LICENSE-WEIGHTS, NOTICE.md

oracle/_ref/ is listed in .gitignore (reference sources and the CC BY-NC
checkpoint are never committed) but not in .gpurunignore, so it travels to the
GPU box with the working tree, exactly like ginfinity_b200/data/ and the built
libgfx.so.  Nothing under ginfinity_b200/ imports it: only tests/,
__graft_entry__.smoke() and bench.py's reference legs do (oracle/ref_loader.py).
A manifest with the SHA-256 of every staged file is written next to the copy so
that a test can check the staged package is byte-identical to what was staged.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEST = HERE / "_ref"
DEFAULT_SOURCE = Path("/root/reference")
EXTRA_FILES = ("LICENSE", "LICENSE-WEIGHTS", "NOTICE.md", "pyproject.toml")


def _sha256(path: Path) -> str:
    return hashlib.sha256(path.read_bytes()).hexdigest()


def stage(source=DEFAULT_SOURCE, dest=DEST) -> Path:
    source, dest = Path(source), Path(dest)
    package = source / "src" / "ginfinity"
    if not (package / "api.py").is_file():
        raise FileNotFoundError(f"{source} is not a GINFINITY checkout")
    if dest.exists():
        shutil.rmtree(dest)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache")
    shutil.copytree(package, dest / "src" / "ginfinity", ignore=ignore)
    (dest / "tests").mkdir(parents=True)
    for path in sorted((source / "tests").iterdir()):
        if path.suffix in (".py", ".tsv"):
            shutil.copyfile(path, dest / "tests" / path.name)
    for name in EXTRA_FILES:
        if (source / name).is_file():
            shutil.copyfile(source / name, dest / name)
    files = {str(p.relative_to(dest)): _sha256(p)
             for p in sorted(dest.rglob("*")) if p.is_file()}
    (dest / "STAGED_MANIFEST.json").write_text(json.dumps(
        {"source": str(source), "files": files}, indent=1) + "\n")
    return dest


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_SOURCE))
