"""Wire-compatible embedding server backed by the GPU recompute stage.

Speaks the msgpack-over-ZMQ REP protocol of the reference's HNSW embedding server
(packages/leann-backend-hnsw/leann_backend_hnsw/hnsw_embedding_server.py:98-340), so an UNMODIFIED LEANN process — the
faiss fork's ZmqDistanceComputer (faiss/impl/HNSW_zmq.cpp:256-420, 579-654) or BaseSearcher.compute_query_embedding
(leann-core/src/leann/searcher_base.py:130-160) — can use the B200 encoder as its embedding server:

  request                              reply
  ["__QUERY_MODEL__"]                  [model_name]
  [str, str, ...]                      embeddings.tolist()                       (texts -> embeddings)
  [[ids], [query_vector]]              [[distances]]  single-float, 1e9 for ids without a passage
  [[ids]]  or  [ids]                   [[n, dim], flat_floats]  single-float, zero rows for ids without a passage

The DiskANN server's variant (leann_backend_diskann/diskann_embedding_server.py:120-210, schema third_party/embedding.proto:5-13)
is provided too: a protobuf NodeEmbeddingRequest {repeated uint32 node_ids = 1} is answered with a NodeEmbeddingResponse
{bytes embeddings_data = 1; repeated int32 dimensions = 2; repeated uint32 missing_ids = 3}; texts arrive as msgpack.  The two
messages are encoded / decoded by the ~40 lines of proto3 wire format below (no generated code, no protobuf runtime).

The request handler is a pure function over two callables (ids -> embeddings, texts -> embeddings), so the protocol
logic is tested without a GPU; `serve()` binds it to a capi.Index and a REP socket.  This is an adapter for the stock
CPU traversal — the product path (lb2_search) never goes through it.
"""
from __future__ import annotations

import logging
import threading
from typing import Callable, Optional, Sequence

import numpy as np

logger = logging.getLogger(__name__)
LARGE_DISTANCE = 1e9  # hnsw_embedding_server.py:183


def handle_request(request, *, encode_ids: Callable[[np.ndarray], np.ndarray],
                   encode_texts: Optional[Callable[[Sequence[str]], np.ndarray]], n_passages: int, dim: int,
                   distance_metric: str, model_name: str):
    """Returns (payload, use_single_float) for one decoded msgpack request."""
    if isinstance(request, list) and len(request) == 1 and request[0] == "__QUERY_MODEL__":
        return [model_name], False
    if isinstance(request, list) and request and all(isinstance(t, str) for t in request):  # :134-145
        if encode_texts is None:
            raise RuntimeError("text requests need a tokenizer for the embedding model")
        return np.asarray(encode_texts(request), np.float32).tolist(), False
    if isinstance(request, list) and len(request) == 2 and isinstance(request[0], list) and isinstance(request[1], list):  # :147-211
        node_ids = request[0]
        if len(node_ids) == 1 and isinstance(node_ids[0], list):
            node_ids = node_ids[0]
        q = np.asarray(request[1], np.float32)
        out = [LARGE_DISTANCE] * len(node_ids)
        found = [i for i, nid in enumerate(node_ids) if isinstance(nid, int) and 0 <= nid < n_passages]
        if found:
            E = np.asarray(encode_ids(np.asarray([node_ids[i] for i in found], np.int64)), np.float32)
            part = np.sum(np.square(E - q.reshape(1, -1)), axis=1) if distance_metric == "l2" else -np.dot(E, q)
            for pos, v in zip(found, part.flatten().tolist()):
                out[pos] = float(v)
        return [out], True
    # embedding-by-id fetch (:213-284)
    if isinstance(request, list) and len(request) == 1 and isinstance(request[0], list):
        node_ids = request[0]
    elif isinstance(request, list):
        node_ids = request
    else:
        node_ids = []
    dims = [len(node_ids), dim]
    flat = np.zeros((len(node_ids), dim), np.float32)
    found = [i for i, nid in enumerate(node_ids) if isinstance(nid, int) and 0 <= nid < n_passages]
    if found:
        E = np.asarray(encode_ids(np.asarray([node_ids[i] for i in found], np.int64)), np.float32)
        if np.isnan(E).any() or np.isinf(E).any():  # :257-262
            return [[0, dim], []], True
        flat[found] = E
    return [dims, flat.reshape(-1).tolist()], True


# ------------------------------------------------------------------------------------------------ proto3 wire format
def _read_varint(buf: bytes, pos: int) -> tuple[int, int]:
    shift = val = 0
    while True:
        if pos >= len(buf) or shift > 63:
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def decode_node_embedding_request(buf: bytes) -> list[int]:
    """NodeEmbeddingRequest.node_ids (field 1: packed, or one varint per occurrence); unknown fields are an error so that
    msgpack text requests fall through to the msgpack branch like in the reference (ParseFromString raising)."""
    ids, pos = [], 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if field != 1 or wt not in (0, 2):
            raise ValueError("not a NodeEmbeddingRequest")
        if wt == 0:
            v, pos = _read_varint(buf, pos)
            ids.append(v)
        else:
            n, pos = _read_varint(buf, pos)
            end = pos + n
            if end > len(buf):
                raise ValueError("truncated packed field")
            while pos < end:
                v, pos = _read_varint(buf, pos)
                ids.append(v)
            if pos != end:
                raise ValueError("bad packed field")
    return ids


def encode_node_embedding_response(emb: np.ndarray, missing: Sequence[int] = ()) -> bytes:
    emb = np.ascontiguousarray(emb, np.float32)
    data = emb.tobytes()
    out = bytearray()
    if data:
        out += b"\x0a" + _varint(len(data)) + data                       # field 1, length-delimited
    dims = b"".join(_varint(int(d)) for d in emb.shape)
    out += b"\x12" + _varint(len(dims)) + dims                            # field 2, packed int32
    if missing:
        ms = b"".join(_varint(int(m)) for m in missing)
        out += b"\x1a" + _varint(len(ms)) + ms                            # field 3, packed uint32
    return bytes(out)


def handle_diskann_request(raw: bytes, *, encode_ids, encode_texts, n_passages: int) -> bytes:
    """diskann_embedding_server.py:124-205: protobuf id request first, msgpack text request as the fallback.  Unknown ids
    raise like the reference's passage lookup (the REP loop turns that into its error reply)."""
    try:
        ids = decode_node_embedding_request(raw)
        if not ids:
            raise ValueError("empty node_ids")
    except ValueError as proto_err:
        import msgpack
        try:
            req = msgpack.unpackb(raw)
        except Exception as e:
            raise RuntimeError(f"Both protobuf and msgpack parsing failed! Protobuf: {proto_err}, Msgpack: {e}") from e
        if not (isinstance(req, list) and all(isinstance(t, str) for t in req)):
            raise RuntimeError(f"Both protobuf and msgpack parsing failed! Protobuf: {proto_err}, Msgpack: not a text request")
        if encode_texts is None:
            raise RuntimeError("text requests need a tokenizer for the embedding model")
        return msgpack.packb(np.asarray(encode_texts(req), np.float32).tolist())
    bad = [i for i in ids if not 0 <= i < n_passages]
    if bad:
        raise KeyError(f"Passage ID not found: {bad[0]}")
    return encode_node_embedding_response(encode_ids(np.asarray(ids, np.int64)))


def serve(index, port: int, distance_metric: str = "mips", model_name: str = "", tokenizer=None, max_len: int = 256,
          shutdown: Optional[threading.Event] = None, ready: Optional[threading.Event] = None, protocol: str = "hnsw") -> None:
    """Blocking REP loop (hnsw_embedding_server.py:105-123: 1 s receive timeout so that `shutdown` is honoured).
    `index`: a capi.Index / capi.DiskannIndex with passages and encoder attached."""
    import msgpack
    import zmq

    def encode_texts(texts):
        if tokenizer is None:
            raise RuntimeError("no tokenizer")
        toks, offs = [], [0]
        for t in texts:
            ids = tokenizer(t, truncation=True, max_length=max_len)["input_ids"]
            toks.extend(ids)
            offs.append(len(toks))
        return index.encode_tokens(np.asarray(toks, np.uint16), np.asarray(offs, np.uint64))

    shutdown = shutdown or threading.Event()
    ctx = zmq.Context()
    sock = ctx.socket(zmq.REP)
    sock.bind(f"tcp://*:{port}")
    sock.setsockopt(zmq.RCVTIMEO, 1000)
    sock.setsockopt(zmq.SNDTIMEO, 1000)
    sock.setsockopt(zmq.LINGER, 0)
    if ready is not None:
        ready.set()
    try:
        while not shutdown.is_set():
            try:
                raw = sock.recv()
            except zmq.Again:
                continue
            if protocol == "diskann":
                try:
                    reply = handle_diskann_request(raw, encode_ids=index.encode_ids,
                                                   encode_texts=encode_texts if tokenizer is not None else None,
                                                   n_passages=int(index.info.ntotal))
                except Exception as e:  # diskann_embedding_server.py:292-312: an empty response keeps the REQ side alive
                    logger.error(f"embedding server request failed: {e}")
                    reply = encode_node_embedding_response(np.zeros((0, int(index.info.d)), np.float32))
                sock.send(reply)
                continue
            try:
                payload, single = handle_request(msgpack.unpackb(raw), encode_ids=index.encode_ids,
                                                 encode_texts=encode_texts if tokenizer is not None else None,
                                                 n_passages=int(index.info.ntotal), dim=int(index.info.d),
                                                 distance_metric=distance_metric, model_name=model_name)
            except Exception as e:  # shape-correct empty reply like the reference's fallback (:296-330)
                logger.error(f"embedding server request failed: {e}")
                payload, single = [[0, int(index.info.d)], []], True
            sock.send(msgpack.packb(payload, use_single_float=single))
    finally:
        sock.close()
        ctx.term()
