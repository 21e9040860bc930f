"""Command line over the encoder path: `python -m ginfinity_b200 <command>`.

Host glue only: the sub-commands, flags, NPZ layout and manifest keys
are those of the reference's `ginfinity` command (src/ginfinity/cli.py:69-197, 226-300), so that
BASELINE config 1 -- `ginfinity embed` over tests/rouskin_sample_6k.tsv -- runs here unchanged
apart from the device: this build has no CPU path, `--device` defaults to "cuda".

    info               verified model metadata as JSON
    alignment-config   scoring parameters for the separate aligner
    embed              delimited RNA table -> <output>.npz + manifest
    build-graphs       delimited RNA table -> graph shard (safetensors + JSON sidecar)
    embed-graphs       graph shard -> <output>.npz + manifest
"""
from __future__ import annotations

import argparse
import hashlib
import json
import platform
import sys
import time
from pathlib import Path

import numpy as np

from . import __version__
from .graph import GraphBuilder, graph_metadata_path, load_graph_shard, save_graph_shard
from .records import read_rna_table

_DTYPES = ("float16", "float32", "float64")


def _file_sha256(path: Path) -> str:
    digest = hashlib.sha256()
    with path.open("rb") as handle:
        for block in iter(lambda: handle.read(1 << 22), b""):
            digest.update(block)
    return digest.hexdigest()


def _encoder(args):
    from .encoder import Ginfinity      # needs libgfx.so and a GPU: imported on demand
    return Ginfinity.load(device=args.device,
                          allow_nondeterministic_cuda=args.allow_nondeterministic_cuda,
                          full_precision=args.full_precision)


def _records(args):
    start, end = (None, None) if args.no_slices else (args.start_column, args.end_column)
    return read_rna_table(args.input, identifier_column=args.id_column,
                          sequence_column=args.sequence_column,
                          structure_column=args.structure_column,
                          start_column=start, end_column=end, delimiter=args.delimiter)


def _context(args) -> tuple[bool, int]:
    """(keep_paired_neighbours, context_hops); giving --context-hops implies keeping the partners
    (cli.py:46-55)."""
    if args.context_hops is not None:
        return True, int(args.context_hops)
    return bool(args.keep_paired_neighbours), 1


def _write_archive(path: Path, names, arrays) -> None:
    """The reference's output format (numpy.savez_compressed, cli.py:86,159), its members
    deflated on every host core (ginfinity_b200/npz.py): same member bytes, a fraction of the
    time -- the archive, not the encode, was the wall time of `embed`."""
    from .npz import write_npz_compressed
    write_npz_compressed(path, names, arrays)


def _write_manifest(path: Path, encoder, body: dict, started: float) -> None:
    meta = encoder.info()
    manifest = {"status": "complete", "ginfinity_version": __version__,
                "model_version": meta.get("model_version"),
                "checkpoint_sha256": meta.get("checkpoint_sha256")}
    manifest.update(body)
    manifest["elapsed_seconds"] = time.time() - started
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(json.dumps(manifest, indent=2) + "\n")


def cmd_info(args) -> int:
    from .encoder import Ginfinity
    print(json.dumps(Ginfinity.load(device=args.device).info(), indent=2))
    return 0


def cmd_alignment_config(args) -> int:
    from .encoder import default_alignment_parameters
    text = json.dumps({"scoring_parameters": default_alignment_parameters()}, indent=2) + "\n"
    if args.output is None:
        sys.stdout.write(text)
    else:
        args.output.parent.mkdir(parents=True, exist_ok=True)
        args.output.write_text(text)
    return 0


def cmd_embed(args) -> int:
    import torch
    started = time.time()
    records = _records(args)
    encoder = _encoder(args)
    keep, hops = _context(args)
    arrays = encoder.encode_many(records, max_batch_nodes=args.max_batch_nodes,
                                 max_batch_edges=args.max_batch_edges,
                                 keep_paired_neighbours=keep, context_hops=hops,
                                 embedding_dtype=args.embedding_dtype)
    _write_archive(args.output, (r.identifier for r in records), arrays)
    entries = []
    for record, value in zip(records, arrays):
        entry = {"identifier": record.identifier, "length": record.length,
                 "core_length": int(value.shape[0]), "shape": list(value.shape)}
        if record.start is not None and record.end is not None:
            entry["start"], entry["end"] = record.start, record.end
        entries.append(entry)
    manifest_path = args.manifest or args.output.with_suffix(".manifest.json")
    _write_manifest(manifest_path, encoder, {
        "input": str(args.input), "input_sha256": _file_sha256(args.input),
        "output": str(args.output), "output_sha256": _file_sha256(args.output),
        "device": args.device, "python": platform.python_version(),
        "numpy": np.__version__, "torch": torch.__version__, "records": entries}, started)
    print(json.dumps({"output": str(args.output), "manifest": str(manifest_path),
                      "records": len(records)}))
    return 0


def cmd_build_graphs(args) -> int:
    started = time.time()
    keep, hops = _context(args)
    shard = GraphBuilder(keep_paired_neighbours=keep, context_hops=hops).build_shard(_records(args))
    metadata_path = args.metadata or graph_metadata_path(args.output)
    save_graph_shard(shard, args.output, metadata_path=metadata_path, checksum=args.checksum)
    print(json.dumps({"output": str(args.output), "metadata": str(metadata_path),
                      "records": shard.record_count, "nodes": shard.node_count,
                      "edges": shard.edge_count, "graph_spec_sha256": shard.spec.sha256,
                      "checksum": args.checksum, "elapsed_seconds": time.time() - started}))
    return 0


def cmd_embed_graphs(args) -> int:
    started = time.time()
    encoder = _encoder(args)
    shard = load_graph_shard(args.input, metadata_path=args.metadata,
                             expected_spec=encoder.graph_spec,
                             verify_checksum=args.verify_checksum,
                             validation="full" if args.full_validation else "metadata")
    arrays = encoder.encode_graphs(shard, max_batch_nodes=args.max_batch_nodes,
                                   max_batch_edges=args.max_batch_edges,
                                   embedding_dtype=args.embedding_dtype)
    _write_archive(args.output, shard.identifiers, arrays)
    entries = [{"identifier": name, "length": len(sequence), "node_count": int(nodes),
                "core_length": int(core), "shape": list(value.shape)}
               for name, sequence, nodes, core, value
               in zip(shard.identifiers, shard.sequences, shard.lengths, shard.core_counts, arrays)]
    body = {"graph_spec_sha256": shard.spec.sha256, "input": str(args.input),
            "input_metadata": str(args.metadata or graph_metadata_path(args.input)),
            "output": str(args.output), "device": args.device, "records": entries}
    if args.checksum:
        body["output_sha256"] = _file_sha256(args.output)
    manifest_path = args.manifest or args.output.with_suffix(".manifest.json")
    _write_manifest(manifest_path, encoder, body, started)
    print(json.dumps({"output": str(args.output), "manifest": str(manifest_path),
                      "records": shard.record_count}))
    return 0


def _table_options(sub) -> None:
    sub.add_argument("--id-column", default="transcript_id")
    sub.add_argument("--sequence-column", default="sequence")
    sub.add_argument("--structure-column", default="secondary_structure")
    sub.add_argument("--start-column", default="start",
                     help="optional column of 0-based half-open window starts")
    sub.add_argument("--end-column", default="end",
                     help="optional column of 0-based half-open window ends")
    sub.add_argument("--no-slices", action="store_true",
                     help="ignore the window columns; encode whole molecules")
    sub.add_argument("--delimiter", default="\t")
    sub.add_argument("--keep-paired-neighbours", action="store_true",
                     help="keep pair partners that fall outside a window as context nodes")
    sub.add_argument("--context-hops", type=int, default=None, metavar="N",
                     help="context depth around those partners (implies "
                          "--keep-paired-neighbours; hop 1 is the partner itself)")


def _model_options(sub) -> None:
    sub.add_argument("--device", default="cuda")
    sub.add_argument("--allow-nondeterministic-cuda", action="store_true",
                     help="accepted for compatibility; this CUDA path is deterministic")
    sub.add_argument("--full-precision", action="store_true",
                     help="float32 activations instead of the default float16")
    sub.add_argument("--max-batch-nodes", type=int, default=60_000)
    sub.add_argument("--max-batch-edges", type=int, default=300_000)
    sub.add_argument("--embedding-dtype", choices=_DTYPES, default="float16",
                     help="dtype of the returned per-nucleotide embeddings")


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="ginfinity_b200")
    parser.add_argument("--version", action="version", version=__version__)
    commands = parser.add_subparsers(dest="command", required=True)

    sub = commands.add_parser("info", help="show verified model metadata")
    sub.add_argument("--device", default="cuda")
    sub.set_defaults(run=cmd_info)

    sub = commands.add_parser("alignment-config", help="export parameters for the aligner")
    sub.add_argument("--output", type=Path)
    sub.set_defaults(run=cmd_alignment_config)

    sub = commands.add_parser("embed", help="encode a delimited table of RNA records")
    sub.add_argument("--input", type=Path, required=True)
    sub.add_argument("--output", type=Path, required=True)
    sub.add_argument("--manifest", type=Path)
    _model_options(sub)
    _table_options(sub)
    sub.set_defaults(run=cmd_embed)

    sub = commands.add_parser("build-graphs", help="build a graph shard from an RNA table")
    sub.add_argument("--input", type=Path, required=True)
    sub.add_argument("--output", type=Path, required=True)
    sub.add_argument("--metadata", type=Path)
    sub.add_argument("--checksum", action="store_true")
    _table_options(sub)
    sub.set_defaults(run=cmd_build_graphs)

    sub = commands.add_parser("embed-graphs", help="encode a previously built graph shard")
    sub.add_argument("--input", type=Path, required=True)
    sub.add_argument("--metadata", type=Path)
    sub.add_argument("--output", type=Path, required=True)
    sub.add_argument("--manifest", type=Path)
    _model_options(sub)
    sub.add_argument("--verify-checksum", action="store_true")
    sub.add_argument("--full-validation", action="store_true")
    sub.add_argument("--checksum", action="store_true")
    sub.set_defaults(run=cmd_embed_graphs)
    return parser


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    try:
        return args.run(args)
    except Exception as error:          # same contract as the reference: message + exit status 2
        print(f"ginfinity_b200: {error}", file=sys.stderr)
        return 2


if __name__ == "__main__":
    raise SystemExit(main())
