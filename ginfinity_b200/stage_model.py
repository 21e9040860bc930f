"""Stage a verified GINFINITY model directory into ginfinity_b200/data/.

    python -m ginfinity_b200.stage_model [SOURCE_DIR]

SOURCE_DIR holds the reference's `encoder.pt`, `model.json` and
`alignment.json` (default: /root/reference/src/ginfinity/data).  The weights
are CC BY-NC: the destination is git-ignored and is never committed; it only
travels with the working tree (e.g. to the GPU box).
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

from .weights import load_checkpoint

DEST = Path(__file__).resolve().parent / "data"
DEFAULT_SOURCE = Path("/root/reference/src/ginfinity/data")


def stage(source=DEFAULT_SOURCE, dest=DEST) -> Path:
    source, dest = Path(source), Path(dest)
    load_checkpoint(source)                      # raises if anything is off
    dest.mkdir(parents=True, exist_ok=True)
    for name in ("encoder.pt", "model.json", "alignment.json"):
        if (source / name).is_file():
            shutil.copyfile(source / name, dest / name)
    return dest


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_SOURCE))
