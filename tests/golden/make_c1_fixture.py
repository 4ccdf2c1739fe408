"""BASELINE config C1 as a committed fixture: the reference's own sample document (data/PrideandPrejudice.txt, 772 389 bytes,
used by apps/document_rag.py and tests/test_basic.py upstream) as ONE uint16 token stream.  No WordPiece vocabulary is
available offline, so tokens come from a seeded hashing tokenizer (documented deviation): lower-cased words and
single punctuation marks -> 1000 + crc32(token) % 29000.  Chunks (254 tokens, stride 127 = the reference's 256/128 split
minus [CLS]/[SEP]) and queries are cut from the stream by tests/helpers.py::load_c1.
    python tests/golden/make_c1_fixture.py      (dev container only: reads /root/reference)
"""
import re
import zlib
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/data/PrideandPrejudice.txt")
OUT = Path(__file__).resolve().parent / "c1_pride_tokens.npz"


def main():
    text = SRC.read_text(encoding="utf-8", errors="ignore")
    toks = re.findall(r"[A-Za-z0-9']+|[^\sA-Za-z0-9']", text)
    ids = np.asarray([1000 + zlib.crc32(t.lower().encode()) % 29000 for t in toks], np.uint16)
    np.savez_compressed(OUT, tokens=ids, source_bytes=np.int64(SRC.stat().st_size))
    print(f"{len(toks)} tokens, {len(set(ids.tolist()))} distinct ids -> {OUT.name} ({OUT.stat().st_size/1e3:.0f} KB)")


if __name__ == "__main__":
    main()
