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


def serve(index, port: int, distance_metric: str = "mips", model_name: str = "", tokenizer=None, max_len: int = 256,
          shutdown: Optional[threading.Event] = None, ready: Optional[threading.Event] = None) -> None:
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
