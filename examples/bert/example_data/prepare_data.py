#!/usr/bin/env python3
"""Turn plain-text corpora into the record stores the BERT example reads.

    python prepare_data.py --train wiki.train.tokens --valid wiki.valid.tokens --out ./data [--build-dict]

Writes ``<out>/train.lmdb`` and ``<out>/valid.lmdb`` (one pickled text line per record, keys are the
decimal record index) and, with ``--build-dict``, a WordPiece-less whitespace vocabulary
``<out>/dict.txt`` headed by the BERT specials.  For real pre-training use the published
``bert-base-uncased`` ``vocab.txt`` as ``dict.txt`` instead (30522 types; one token per line).

``--format lmdb`` needs the ``lmdb`` package (not a dependency of the framework itself); ``--format records``
writes the dependency-free ``unicore.data.record_store`` file under the same name (the reader recognises
either); ``auto`` picks LMDB when it is importable.
"""
import argparse
import collections
import os
import pickle


def iter_lines(path, min_chars):
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if len(line) >= min_chars and not line.startswith("="):
                yield line


def write_records(lines, out_path):
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
    from unicore.data.record_store import RecordStoreWriter

    n = 0
    with RecordStoreWriter(out_path) as w:
        for line in lines:
            w.append(line)
            n += 1
    return n


def have_lmdb():
    try:
        import lmdb  # noqa: F401
    except ImportError:
        return False
    return True


def write_store(lines, out_path, map_gb):
    import lmdb

    if os.path.exists(out_path):
        os.remove(out_path)
    env = lmdb.open(out_path, subdir=False, readonly=False, lock=False, readahead=False, meminit=False,
                    max_readers=1, map_size=int(map_gb * (1 << 30)))
    n = 0
    txn = env.begin(write=True)
    for line in lines:
        txn.put(str(n).encode("ascii"), pickle.dumps(line, protocol=pickle.HIGHEST_PROTOCOL))
        n += 1
        if n % 10000 == 0:
            txn.commit()
            txn = env.begin(write=True)
    txn.commit()
    env.close()
    return n


def build_dict(paths, out_path, min_count, min_chars):
    counts = collections.Counter()
    for p in paths:
        for line in iter_lines(p, min_chars):
            counts.update(line.lower().split())
    with open(out_path, "w", encoding="utf-8") as f:
        for special in ("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"):
            f.write(special + "\n")
        for tok, c in counts.most_common():
            if c >= min_count:
                f.write(tok + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", required=True)
    ap.add_argument("--valid", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--min-chars", type=int, default=16)
    ap.add_argument("--map-gb", type=float, default=64)
    ap.add_argument("--build-dict", action="store_true")
    ap.add_argument("--min-count", type=int, default=5)
    ap.add_argument("--format", choices=["auto", "lmdb", "records"], default="auto")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    fmt = a.format if a.format != "auto" else ("lmdb" if have_lmdb() else "records")
    for split, src in (("train", a.train), ("valid", a.valid)):
        dst = os.path.join(a.out, split + ".lmdb")
        lines = iter_lines(src, a.min_chars)
        n = write_store(lines, dst, a.map_gb) if fmt == "lmdb" else write_records(lines, dst)
        print("{}: {} records ({})".format(split, n, fmt))
    if a.build_dict:
        build_dict([a.train], os.path.join(a.out, "dict.txt"), a.min_count, a.min_chars)


if __name__ == "__main__":
    main()
