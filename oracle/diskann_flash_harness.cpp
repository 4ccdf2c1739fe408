// TEST INFRASTRUCTURE (oracle/) — never linked into the product library.
//
// The reference's own DiskANN search loop, compiled from the source where it lies:
//   third_party/DiskANN/src/pq_flash_index.cpp  (PQFlashIndex::load, cached_beam_search :1779-2906, fetch_embeddings_zmq
//   :1579-1710, preprocess_fetched_embeddings :1723-1777) + scratch.cpp, distance.cpp, pq.cpp, logger.cpp, ann_exception.cpp
// behind declaration-only stand-ins for what this container lacks (oracle/stubs: libaio.h, mkl.h, boost/dynamic_bitset.hpp,
// embedding.pb.h = the two proto3 messages with real wire encoding) and with the two out-of-process pieces replaced in
// process:
//   * AlignedFileReader -> PreadFileReader below (synchronous pread of the same 4 KB sectors; same bytes, no libaio)
//   * libzmq            -> the zmq_* functions below: a REQ "socket" whose reply is produced by decoding the
//                          NodeEmbeddingRequest the reference serialised, looking the ids up in a table of embeddings
//                          handed in by the test (the role of diskann_embedding_server.py), and encoding a
//                          NodeEmbeddingResponse the reference parses.
// extern "C" window: dflash_open / dflash_set_embeddings / dflash_search / dflash_close; tests/test_vamana_oracle.py pins
// oracle/vamana_oracle.c (and through it the CUDA path) to this on ids, distances, cmps, hops and I/O counts.
#include <fcntl.h>
#include <unistd.h>
#include <zmq.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "embedding.pb.h"
#include "mkl.h"
#include "percentile_stats.h"
#include "pq_flash_index.h"

// ---------------------------------------------------------------- training-only dependencies of pq.cpp (never reached)
#define LB2_UNREACHABLE(name) { std::fprintf(stderr, "diskann_flash_harness: %s reached\n", name); std::abort(); }
extern "C" {
void cblas_sgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, MKL_INT, MKL_INT, MKL_INT, float, const float*, MKL_INT,
                 const float*, MKL_INT, float, float*, MKL_INT) LB2_UNREACHABLE("cblas_sgemm")
int LAPACKE_sgesdd(int, char, MKL_INT, MKL_INT, float*, MKL_INT, float*, float*, MKL_INT, float*, MKL_INT) LB2_UNREACHABLE("LAPACKE_sgesdd")
}
namespace math_utils {
void compute_closest_centers(float*, size_t, size_t, float*, size_t, size_t, uint32_t*, std::vector<size_t>*, float*)
    LB2_UNREACHABLE("compute_closest_centers")
}
namespace kmeans {
float run_lloyds(float*, size_t, size_t, float*, const size_t, const size_t, std::vector<size_t>*, uint32_t*) LB2_UNREACHABLE("run_lloyds")
void kmeanspp_selecting_pivots(float*, size_t, size_t, float*, size_t) LB2_UNREACHABLE("kmeanspp_selecting_pivots")
}
template <typename T>
void gen_random_slice(const std::string, double, float*&, size_t&, size_t&) LB2_UNREACHABLE("gen_random_slice")
template void gen_random_slice<uint8_t>(const std::string, double, float*&, size_t&, size_t&);
template void gen_random_slice<int8_t>(const std::string, double, float*&, size_t&, size_t&);
template void gen_random_slice<float>(const std::string, double, float*&, size_t&, size_t&);

// ---------------------------------------------------------------- in-process embedding "server"
namespace {
std::mutex g_emb_mu;
const float* g_emb = nullptr;  // [g_emb_n, g_emb_d], owned by the caller
int64_t g_emb_n = 0;
int g_emb_d = 0;
int64_t g_fetch_calls = 0, g_fetch_ids = 0;

struct FakeSocket {
    std::string reply;
    bool has_reply = false;
};
}  // namespace

extern "C" {
void* zmq_ctx_new(void) { return reinterpret_cast<void*>(0x1); }
int zmq_ctx_destroy(void*) { return 0; }
void* zmq_socket(void*, int) { return new FakeSocket(); }
int zmq_close(void* s) { delete static_cast<FakeSocket*>(s); return 0; }
int zmq_setsockopt(void*, int, const void*, size_t) { return 0; }
int zmq_connect(void*, const char*) { return 0; }
int zmq_errno(void) { return 0; }
const char* zmq_strerror(int) { return "in-process zmq stand-in"; }
int zmq_send(void* s, const void* buf, size_t len, int) {
    FakeSocket* sock = static_cast<FakeSocket*>(s);
    protoembedding::NodeEmbeddingRequest req;
    if (!req.ParseFromArray(buf, static_cast<int>(len))) return -1;
    std::lock_guard<std::mutex> lk(g_emb_mu);
    if (!g_emb) return -1;
    const int n = req.node_ids_size();
    std::vector<float> data(static_cast<size_t>(n) * g_emb_d);
    for (int i = 0; i < n; i++) {
        const uint32_t id = req.node_ids(i);
        if (id >= g_emb_n) return -1;
        std::memcpy(&data[static_cast<size_t>(i) * g_emb_d], g_emb + static_cast<size_t>(id) * g_emb_d, sizeof(float) * g_emb_d);
    }
    protoembedding::NodeEmbeddingResponse resp;
    resp.set_embeddings_data(data.data(), data.size() * sizeof(float));
    resp.add_dimensions(n);
    resp.add_dimensions(g_emb_d);
    resp.SerializeToString(&sock->reply);
    sock->has_reply = true;
    g_fetch_calls++;
    g_fetch_ids += n;
    return static_cast<int>(len);
}
// zmq_msg_t is an opaque 64-byte blob: the first pointer-sized slot holds our std::string*
int zmq_msg_init(zmq_msg_t* m) { std::memset(m, 0, sizeof(*m)); return 0; }
int zmq_msg_recv(zmq_msg_t* m, void* s, int) {
    FakeSocket* sock = static_cast<FakeSocket*>(s);
    if (!sock->has_reply) return -1;
    std::string* p = new std::string(std::move(sock->reply));
    sock->has_reply = false;
    std::memcpy(m, &p, sizeof(p));
    return static_cast<int>(p->size());
}
static std::string* msg_str(zmq_msg_t* m) { std::string* p; std::memcpy(&p, m, sizeof(p)); return p; }
void* zmq_msg_data(zmq_msg_t* m) { return msg_str(m) ? const_cast<char*>(msg_str(m)->data()) : nullptr; }
size_t zmq_msg_size(const zmq_msg_t* m) { return msg_str(const_cast<zmq_msg_t*>(m)) ? msg_str(const_cast<zmq_msg_t*>(m))->size() : 0; }
int zmq_msg_close(zmq_msg_t* m) { delete msg_str(m); std::memset(m, 0, sizeof(*m)); return 0; }
}

// ---------------------------------------------------------------- synchronous stand-in for LinuxAlignedFileReader
namespace {
class PreadFileReader : public AlignedFileReader {
    int fd_ = -1;
    IOContext ctx_ = nullptr;

  public:
    IOContext& get_ctx() override { return ctx_; }
    void register_thread() override {}
    void deregister_thread() override {}
    void deregister_all_threads() override {}
    void open(const std::string& fname) override {
        fd_ = ::open(fname.c_str(), O_RDONLY);
        if (fd_ < 0) { std::fprintf(stderr, "diskann_flash_harness: cannot open %s\n", fname.c_str()); }
    }
    void close() override { if (fd_ >= 0) ::close(fd_); fd_ = -1; }
    void read(std::vector<AlignedRead>& reqs, IOContext&, bool) override {
        for (auto& r : reqs) {
            size_t done = 0;
            while (done < r.len) {
                const ssize_t n = ::pread(fd_, static_cast<char*>(r.buf) + done, r.len - done, static_cast<off_t>(r.offset + done));
                if (n <= 0) { std::memset(static_cast<char*>(r.buf) + done, 0, r.len - done); break; }  // past EOF: zeros
                done += static_cast<size_t>(n);
            }
        }
    }
    ~PreadFileReader() override { close(); }
};

struct Flash {
    std::shared_ptr<AlignedFileReader> reader, graph_reader;
    std::unique_ptr<diskann::PQFlashIndex<float>> index;
};
}  // namespace

extern "C" {

// metric: 0 l2, 1 mips, 2 cosine (include/leann_b200.h LB2_METRIC_*)
void* dflash_open(const char* index_prefix, const char* pq_prefix, const char* partition_prefix, int metric, int nthreads) {
    try {
        Flash* f = new Flash();
        f->reader = std::make_shared<PreadFileReader>();
        f->graph_reader = std::make_shared<PreadFileReader>();
        const diskann::Metric m = metric == 1 ? diskann::Metric::INNER_PRODUCT : metric == 2 ? diskann::Metric::COSINE : diskann::Metric::L2;
        f->index.reset(new diskann::PQFlashIndex<float>(f->reader, f->graph_reader, m));
        const int rc = f->index->load(static_cast<uint32_t>(nthreads > 0 ? nthreads : 1), index_prefix, /*zmq_port=*/5555,
                                      pq_prefix ? pq_prefix : "", partition_prefix ? partition_prefix : "");  // the binding passes c_str()s, never null
        if (rc != 0) { delete f; return nullptr; }
        return f;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "diskann_flash_harness: load failed: %s\n", e.what());
        return nullptr;
    }
}

void dflash_close(void* h) { delete static_cast<Flash*>(h); }

void dflash_set_embeddings(const float* emb, int64_t n, int d) {
    std::lock_guard<std::mutex> lk(g_emb_mu);
    g_emb = emb; g_emb_n = n; g_emb_d = d;
}

void dflash_fetch_counters(int64_t* calls, int64_t* ids) { *calls = g_fetch_calls; *ids = g_fetch_ids; }

// StaticDiskIndex::batch_search (python/src/static_disk_index.cpp:88-118), one query at a time: stats per query
int dflash_search(void* h, const float* q, int64_t nq, int dim, int64_t k, int64_t L, int64_t beam, int deferred_fetch,
                  int skip_search_reorder, uint32_t io_limit, uint64_t* ids, float* dists, uint32_t* cmps, uint32_t* hops,
                  uint32_t* ios) {
    Flash* f = static_cast<Flash*>(h);
    try {
        for (int64_t i = 0; i < nq; i++) {
            diskann::QueryStats st;
            f->index->cached_beam_search(q + i * dim, static_cast<uint64_t>(k), static_cast<uint64_t>(L), ids + i * k, dists + i * k,
                                         static_cast<uint64_t>(beam), io_limit, /*use_reorder_data=*/false, &st, deferred_fetch != 0,
                                         skip_search_reorder != 0, /*recompute_beighbor_embeddings=*/false, /*dedup_node_dis=*/false,
                                         /*prune_ratio=*/0.f, /*batch_recompute=*/false, /*global_pruning=*/true);
            if (cmps) cmps[i] = st.n_cmps;
            if (hops) hops[i] = st.n_hops;
            if (ios) ios[i] = st.n_ios;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "diskann_flash_harness: search failed: %s\n", e.what());
        return -1;
    }
    return 0;
}

}  // extern "C"
