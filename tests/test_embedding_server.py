"""Wire protocol of the embedding-server shim (msgpack over ZMQ REP), host logic only: a fake encoder stands in for the GPU.
Request/reply shapes follow hnsw_embedding_server.py:128-284 and the C++ client's expectations (HNSW_zmq.cpp:256-420, 579-654)."""
import threading
import types

import numpy as np
import pytest

from leann_b200.embedding_server import (LARGE_DISTANCE, decode_node_embedding_request, encode_node_embedding_response,
                                         handle_diskann_request, handle_request, serve)

DIM, N = 8, 20
TABLE = np.random.default_rng(0).standard_normal((N, DIM)).astype(np.float32)


def enc_ids(ids):
    return TABLE[np.asarray(ids)]


def call(req, metric="mips", texts=None):
    return handle_request(req, encode_ids=enc_ids, encode_texts=texts, n_passages=N, dim=DIM, distance_metric=metric,
                          model_name="m")


def test_request_kinds():
    assert call(["__QUERY_MODEL__"]) == (["m"], False)
    q = TABLE[3].tolist()
    (out,), single = call([[1, 2, 999, 3], q])
    assert single and out[2] == LARGE_DISTANCE and np.allclose([out[0], out[1], out[3]], -(TABLE[[1, 2, 3]] @ TABLE[3]), atol=1e-6)
    (out2,), _ = call([[[1, 2]], q], metric="l2")  # nested [[ids]] accepted (server :158-160)
    assert np.allclose(out2, ((TABLE[[1, 2]] - TABLE[3]) ** 2).sum(1), atol=1e-5)
    (dims, flat), single = call([[4, 25, 5]])
    assert dims == [3, DIM] and single
    got = np.asarray(flat, np.float32).reshape(3, DIM)
    assert np.array_equal(got[0], TABLE[4]) and not got[1].any() and np.array_equal(got[2], TABLE[5])
    (dims, flat), _ = call([7, 8])  # bare id list
    assert dims == [2, DIM] and np.array_equal(np.asarray(flat, np.float32).reshape(2, DIM), TABLE[[7, 8]])
    emb, single = call(["a", "bb"], texts=lambda ts: np.stack([np.full(DIM, len(t), np.float32) for t in ts]))
    assert not single and emb == [[1.0] * DIM, [2.0] * DIM]
    with pytest.raises(RuntimeError, match="tokenizer"):
        call(["a"])


def test_zmq_round_trip_like_the_cpp_client():
    zmq = pytest.importorskip("zmq")
    msgpack = pytest.importorskip("msgpack")
    fake = types.SimpleNamespace(encode_ids=enc_ids, info=types.SimpleNamespace(ntotal=N, d=DIM))
    stop, ready = threading.Event(), threading.Event()
    port = 5990 + (np.random.default_rng().integers(0, 400))
    th = threading.Thread(target=serve, args=(fake, int(port)), kwargs=dict(distance_metric="mips", model_name="m", shutdown=stop, ready=ready), daemon=True)
    th.start()
    assert ready.wait(10)
    ctx = zmq.Context()
    s = ctx.socket(zmq.REQ)
    s.setsockopt(zmq.RCVTIMEO, 5000)
    s.setsockopt(zmq.LINGER, 0)
    s.connect(f"tcp://127.0.0.1:{port}")
    try:
        s.send(msgpack.packb([[1, 2, 3], TABLE[0].tolist()]))  # distances_batch request (HNSW_zmq.cpp:579-654)
        (dist,) = msgpack.unpackb(s.recv())
        assert np.allclose(dist, -(TABLE[[1, 2, 3]] @ TABLE[0]), atol=1e-6)
        s.send(msgpack.packb([[5, 6]]))                          # fetch_embeddings_zmq request (:256-420)
        dims, flat = msgpack.unpackb(s.recv())
        assert dims == [2, DIM] and np.array_equal(np.asarray(flat, np.float32).reshape(2, DIM), TABLE[[5, 6]])
        s.send(msgpack.packb(["text without tokenizer"]))        # failure -> shape-correct empty reply, server stays up
        assert msgpack.unpackb(s.recv()) == [[0, DIM], []]
        s.send(msgpack.packb(["__QUERY_MODEL__"]))
        assert msgpack.unpackb(s.recv()) == ["m"]
    finally:
        stop.set()
        s.close()
        ctx.term()
        th.join(5)


def _proto_classes():
    """The two messages of third_party/embedding.proto built with the protobuf runtime's dynamic descriptors (no generated
    code): the independent codec the hand-written wire format is checked against."""
    pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="embedding_test.proto", package="protoembedding_test", syntax="proto3")
    req = fd.message_type.add(name="NodeEmbeddingRequest")
    req.field.add(name="node_ids", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_UINT32,
                  label=descriptor_pb2.FieldDescriptorProto.LABEL_REPEATED)
    resp = fd.message_type.add(name="NodeEmbeddingResponse")
    resp.field.add(name="embeddings_data", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_BYTES,
                   label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    resp.field.add(name="dimensions", number=2, type=descriptor_pb2.FieldDescriptorProto.TYPE_INT32,
                   label=descriptor_pb2.FieldDescriptorProto.LABEL_REPEATED)
    resp.field.add(name="missing_ids", number=3, type=descriptor_pb2.FieldDescriptorProto.TYPE_UINT32,
                   label=descriptor_pb2.FieldDescriptorProto.LABEL_REPEATED)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = message_factory.GetMessageClass
    return get(pool.FindMessageTypeByName("protoembedding_test.NodeEmbeddingRequest")), \
        get(pool.FindMessageTypeByName("protoembedding_test.NodeEmbeddingResponse"))


def test_diskann_protobuf_variant_against_the_protobuf_runtime():
    Req, Resp = _proto_classes()
    for ids in ([3], [0, 1, 19], list(range(0, 20, 3)), [2 ** 31 + 5, 7][1:]):
        raw = Req(node_ids=ids).SerializeToString()
        assert decode_node_embedding_request(raw) == ids
        out = Resp()
        out.ParseFromString(handle_diskann_request(raw, encode_ids=enc_ids, encode_texts=None, n_passages=N))
        assert list(out.dimensions) == [len(ids), DIM] and not out.missing_ids
        assert np.array_equal(np.frombuffer(out.embeddings_data, np.float32).reshape(len(ids), DIM), TABLE[ids])
    r = Resp()
    r.ParseFromString(encode_node_embedding_response(np.zeros((0, DIM), np.float32), missing=[4, 300]))
    assert list(r.dimensions) == [0, DIM] and list(r.missing_ids) == [4, 300] and r.embeddings_data == b""
    with pytest.raises(KeyError):
        handle_diskann_request(Req(node_ids=[1, 999]).SerializeToString(), encode_ids=enc_ids, encode_texts=None, n_passages=N)
    # msgpack text fallback (BaseSearcher.compute_query_embedding), and garbage
    msgpack = pytest.importorskip("msgpack")
    emb = msgpack.unpackb(handle_diskann_request(msgpack.packb(["ab", "c"]), encode_ids=enc_ids, n_passages=N,
                                                 encode_texts=lambda ts: np.stack([np.full(DIM, len(t), np.float32) for t in ts])))
    assert emb == [[2.0] * DIM, [1.0] * DIM]
    with pytest.raises(RuntimeError, match="Both protobuf and msgpack parsing failed"):
        handle_diskann_request(b"\xff\xff\xff", encode_ids=enc_ids, encode_texts=None, n_passages=N)
