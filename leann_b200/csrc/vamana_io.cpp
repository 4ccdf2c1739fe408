// Reader for the files PQFlashIndex::load opens (DiskANN/src/pq_flash_index.cpp:887-911, 1017-1467; all
// citations under /root/reference/packages/leann-backend-diskann/third_party/DiskANN/):
//   <p>_pq_pivots.bin            FixedChunkPQTable::load_pq_centroid_bin, src/pq.cpp:49-168
//   <p>_pq_compressed.bin        bin<u8>[npts, n_chunks], pq_flash_index.cpp:1056-1063
//   <p>_disk.index               metadata sector :1292-1362, node layout :118-141  (not opened in partition mode)
//   <p>_disk.index_medoids.bin / _centroids.bin / _max_base_norm.bin   :1369-1452
//   <pp>_partition.bin, <pp>_disk_graph.index   read_partition_info :916-948, load_graph_index :951-1008,
//                                adjacency-only sectors :2248-2290, 2509-2564
// Everything is validated (ids < npts, degrees <= R, codes complete) so the kernels index without bounds checks.
// Not supported (the reference's LEANN backend never produces them): OPQ rotation matrix, disk-PQ
// (_disk.index_pq_pivots.bin), filter labels / dummy points, frozen points, reorder data.
#include "vamana_io.h"

#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>

namespace lb2 {
namespace {

constexpr uint64_t SECTOR = 4096;

bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

struct File {
    FILE* f = nullptr;
    std::string path;
    std::string* err;
    explicit File(const std::string& p, std::string* e) : path(p), err(e) { f = fopen(p.c_str(), "rb"); if (!f) *err = p + ": cannot open"; }
    ~File() { if (f) fclose(f); }
    bool ok() const { return f != nullptr; }
    bool at(uint64_t off) { if (fseeko(f, (off_t)off, SEEK_SET) != 0) { *err = path + ": seek failed"; return false; } return true; }
    bool raw(void* dst, size_t n) { if (n && fread(dst, 1, n, f) != n) { *err = path + ": unexpected end of file"; return false; } return true; }
    // DiskANN "bin": i32 rows, i32 cols, row-major payload (include/utils.h load_bin)
    template <class T>
    bool bin(uint64_t off, std::vector<T>* v, int64_t* rows, int64_t* cols, uint64_t max_elems) {
        int32_t r = 0, c = 0;
        if (!at(off) || !raw(&r, 4) || !raw(&c, 4)) return false;
        if (r < 0 || c < 0 || (uint64_t)r * (uint64_t)c > max_elems) { *err = path + ": implausible bin header"; return false; }
        v->resize((size_t)r * (size_t)c);
        *rows = r; *cols = c;
        return raw(v->data(), v->size() * sizeof(T));
    }
};

}  // namespace

// <p>_pq_compressed.bin + <p>_pq_pivots.bin (also what HNSW::load_pq_pruning_data reads for PQ-guided pruning,
// leann-backend-hnsw/third_party/faiss/faiss/impl/HNSW_search.cpp:253-297, impl/pq.cpp:41-138)
bool read_pq_files(const std::string& pivots, const std::string& compressed, PqHost* o, std::string* err) {
    int64_t r = 0, c = 0;
    {   // ---- PQ codes
        File f(compressed, err);
        if (!f.ok()) return false;
        if (!f.bin<uint8_t>(0, &o->codes, &r, &c, 1ull << 40)) return false;
        o->n = r; o->n_chunks = (int)c;
        if (o->n <= 0 || o->n >= (1ll << 31) || o->n_chunks <= 0 || o->n_chunks > 512) { *err = compressed + ": bad shape"; return false; }
    }
    {   // ---- PQ pivots
        File f(pivots, err);
        if (!f.ok()) return false;
        std::vector<uint64_t> offs;
        if (!f.bin<uint64_t>(0, &offs, &r, &c, 16)) return false;
        if (r != 4 && r != 5) { *err = pivots + ": expected 4 or 5 offsets"; return false; }
        std::vector<float> tables;
        if (!f.bin<float>(offs[0], &tables, &r, &c, 256ull * 65536)) return false;
        if (r != 256 || c <= 0) { *err = pivots + ": expected 256 pivots"; return false; }
        o->ndims = (int)c;
        if (!f.bin<float>(offs[1], &o->centroid, &r, &c, 65536)) return false;
        if (r != o->ndims || c != 1) { *err = pivots + ": centroid shape mismatch"; return false; }
        const bool old_type = offs.size() == 5;  // 5-offset files keep the chunk offsets in slot 3 (src/pq.cpp:126-131)
        if (!f.bin<uint32_t>(offs[old_type ? 3 : 2], &o->chunk_offsets, &r, &c, 1024)) return false;
        if (c != 1 || r != o->n_chunks + 1) { *err = pivots + ": chunk offsets do not match the compressed file"; return false; }
        for (int i = 0; i < o->n_chunks; i++)
            if (o->chunk_offsets[i] > o->chunk_offsets[i + 1] || o->chunk_offsets[i + 1] > (uint32_t)o->ndims) { *err = pivots + ": chunk offsets out of range"; return false; }
        o->tables_tr.resize((size_t)o->ndims * 256);  // src/pq.cpp:158-166
        for (int i = 0; i < 256; i++)
            for (int j = 0; j < o->ndims; j++) o->tables_tr[(size_t)j * 256 + i] = tables[(size_t)i * o->ndims + j];
    }
    return true;
}

bool read_diskann_index(const char* index_prefix, const char* partition_prefix, int metric, VamanaHost* o, std::string* err) {
    const std::string p = index_prefix ? index_prefix : "";
    const std::string pp = partition_prefix ? partition_prefix : "";
    if (p.empty()) { *err = "index_prefix is empty"; return false; }
    if (metric < 0 || metric > 2) { *err = "metric must be 0 (l2), 1 (mips) or 2 (cosine)"; return false; }
    o->metric = metric;
    o->partitioned = !pp.empty();
    const std::string pivots = p + "_pq_pivots.bin", compressed = p + "_pq_compressed.bin", disk = p + "_disk.index";
    if (exists(pivots + "_rotation_matrix.bin")) { *err = "OPQ rotation matrix is not supported"; return false; }
    if (!o->partitioned && exists(disk + "_pq_pivots.bin")) { *err = "disk-PQ indexes are not supported"; return false; }
    if (exists(disk + "_labels.txt")) { *err = "filtered (labelled) indexes are not supported"; return false; }
    int64_t r = 0, c = 0;
    {   // ---- PQ codes + pivots
        PqHost pq;
        if (!read_pq_files(pivots, compressed, &pq, err)) return false;
        o->n = pq.n; o->n_chunks = pq.n_chunks; o->data_dim = pq.ndims;
        o->codes.swap(pq.codes); o->tables_tr.swap(pq.tables_tr); o->centroid.swap(pq.centroid); o->chunk_offsets.swap(pq.chunk_offsets);
    }
    const int D = o->data_dim;
    uint64_t medoid_on_file = 0;
    if (!o->partitioned) {  // ---- standard layout: coordinates + adjacency per node
        File f(disk, err);
        if (!f.ok()) return false;
        std::vector<uint64_t> meta;
        if (!f.bin<uint64_t>(0, &meta, &r, &c, 64)) return false;
        if (meta.size() < 8) { *err = disk + ": metadata too short"; return false; }
        const uint64_t nnodes = meta[0], ndims = meta[1], max_node_len = meta[3], nps = meta[4];
        medoid_on_file = meta[2];
        if ((int64_t)nnodes != o->n) { *err = disk + ": point count differs from the PQ file"; return false; }
        if ((int)ndims != D) { *err = disk + ": dimension differs from the PQ pivots"; return false; }
        if (meta[5] != 0) { *err = "indexes with frozen points are not supported"; return false; }
        if (meta[7] != 0) { *err = "indexes with reorder data are not supported"; return false; }
        const uint64_t bpp = (uint64_t)D * 4;
        if (max_node_len < bpp + 8 || max_node_len > (1u << 20)) { *err = disk + ": bad max_node_len"; return false; }
        o->R = (int)((max_node_len - bpp) / 4 - 1);
        const uint64_t spn = (max_node_len + SECTOR - 1) / SECTOR;
        o->nbrs.assign((size_t)o->n * o->R, -1);
        o->coords.resize((size_t)o->n * D);
        std::vector<char> buf((size_t)std::max<uint64_t>(spn, 1) * SECTOR);
        for (int64_t i = 0; i < o->n; i++) {
            const uint64_t sector = 1 + (nps > 0 ? (uint64_t)i / nps : (uint64_t)i * spn);
            const uint64_t in_sector = nps > 0 ? ((uint64_t)i % nps) * max_node_len : 0;
            if (nps == 0 || (uint64_t)i % nps == 0) {
                if (!f.at(sector * SECTOR) || !f.raw(buf.data(), (nps > 0 ? 1 : spn) * SECTOR)) return false;
            }
            const char* node = buf.data() + in_sector;
            memcpy(&o->coords[(size_t)i * D], node, bpp);
            uint32_t nn = 0;
            memcpy(&nn, node + bpp, 4);
            if (nn > (uint32_t)o->R) { *err = disk + ": node degree exceeds max degree"; return false; }
            const uint32_t* nb = reinterpret_cast<const uint32_t*>(node + bpp + 4);
            for (uint32_t m = 0; m < nn; m++) {
                uint32_t v; memcpy(&v, nb + m, 4);
                if ((int64_t)v >= o->n) { *err = disk + ": neighbour id out of range"; return false; }
                o->nbrs[(size_t)i * o->R + m] = (int32_t)v;
            }
            o->n_edges += nn;
        }
    } else {  // ---- partition mode: adjacency-only sectors addressed through id -> partition
        std::vector<std::vector<uint32_t>> parts;
        std::vector<uint32_t> id2p;
        uint64_t C = 0, nparts = 0, nd = 0;
        {
            File f(pp + "_partition.bin", err);
            if (!f.ok()) return false;
            if (!f.raw(&C, 8) || !f.raw(&nparts, 8) || !f.raw(&nd, 8)) return false;
            if ((int64_t)nd != o->n || nparts == 0 || nparts > nd) { *err = pp + "_partition.bin: header does not match the PQ file"; return false; }
            parts.resize(nparts);
            for (uint64_t i = 0; i < nparts; i++) {
                uint32_t sz = 0;
                if (!f.raw(&sz, 4)) return false;
                if (sz > nd) { *err = pp + "_partition.bin: bad partition size"; return false; }
                parts[i].resize(sz);
                if (!f.raw(parts[i].data(), (size_t)sz * 4)) return false;
            }
            id2p.resize(nd);
            if (!f.raw(id2p.data(), (size_t)nd * 4)) return false;
        }
        File g(pp + "_disk_graph.index", err);
        if (!g.ok()) return false;
        int32_t meta_n = 0, meta_dim = 0;
        if (!g.at(0) || !g.raw(&meta_n, 4) || !g.raw(&meta_dim, 4)) return false;
        if (meta_n < 9 || meta_n > 64) { *err = pp + "_disk_graph.index: bad metadata"; return false; }
        std::vector<uint64_t> meta((size_t)meta_n);
        if (!g.raw(meta.data(), meta.size() * 8)) return false;
        const uint64_t dim_in_meta = meta[1], max_node_len = meta[3];
        if (max_node_len <= dim_in_meta * 4 + 4) { *err = pp + "_disk_graph.index: bad max_node_len"; return false; }
        const uint64_t graph_node_len = max_node_len - dim_in_meta * 4;
        o->R = (int)(graph_node_len / 4 - 1);
        o->nbrs.assign((size_t)o->n * o->R, -1);
        std::vector<char> buf(SECTOR);
        std::vector<uint8_t> seen((size_t)o->n, 0);
        for (uint64_t pi = 0; pi < nparts; pi++) {
            if (!g.at((pi + 1) * SECTOR) || !g.raw(buf.data(), SECTOR)) return false;
            for (size_t j = 0; j < parts[pi].size(); j++) {
                const uint32_t id = parts[pi][j];
                if ((int64_t)id >= o->n || id2p[id] != pi) { *err = pp + "_partition.bin: id/partition maps disagree"; return false; }
                const uint64_t off = j * graph_node_len;
                uint32_t nn = 0;
                if (off + 4 > SECTOR) { *err = pp + "_disk_graph.index: node offset out of range"; return false; }
                memcpy(&nn, buf.data() + off, 4);
                if (nn > (uint32_t)o->R || off + 4 + (uint64_t)nn * 4 > SECTOR) { *err = pp + "_disk_graph.index: neighbour data out of range"; return false; }
                for (uint32_t m = 0; m < nn; m++) {
                    uint32_t v; memcpy(&v, buf.data() + off + 4 + 4 * m, 4);
                    if ((int64_t)v >= o->n) { *err = pp + "_disk_graph.index: neighbour id out of range"; return false; }
                    o->nbrs[(size_t)id * o->R + m] = (int32_t)v;
                }
                o->n_edges += nn;
                seen[id] = 1;
            }
        }
        for (int64_t i = 0; i < o->n; i++)
            if (!seen[i]) { *err = pp + "_partition.bin: a node belongs to no partition"; return false; }
    }
    // ---- medoids (+ centroids used to pick among several medoids)
    const std::string medoids_file = disk + "_medoids.bin", centroids_file = disk + "_centroids.bin";
    if (exists(medoids_file)) {
        File f(medoids_file, err);
        if (!f.ok()) return false;
        if (!f.bin<uint32_t>(0, &o->medoids, &r, &c, 1 << 20)) return false;
        if (c != 1 || r < 1) { *err = medoids_file + ": expected an m x 1 vector"; return false; }
    } else {
        if (o->partitioned) { *err = medoids_file + " is required in partition mode"; return false; }
        o->medoids.assign(1, (uint32_t)medoid_on_file);
    }
    for (uint32_t m : o->medoids)
        if ((int64_t)m >= o->n) { *err = "medoid id out of range"; return false; }
    if (o->medoids.size() > 1) {
        if (exists(centroids_file)) {
            File f(centroids_file, err);
            if (!f.ok()) return false;
            if (!f.bin<float>(0, &o->centroid_data, &r, &c, 1ull << 32)) return false;
            if ((size_t)r != o->medoids.size() || c != D) { *err = centroids_file + ": shape mismatch"; return false; }
        } else if (!o->coords.empty()) {  // use_medoids_data_as_centroids
            o->centroid_data.resize(o->medoids.size() * (size_t)D);
            for (size_t m = 0; m < o->medoids.size(); m++)
                memcpy(&o->centroid_data[m * D], &o->coords[(size_t)o->medoids[m] * D], (size_t)D * 4);
        } else { *err = "several medoids but no centroid data"; return false; }
    }
    // ---- max base norm (MIPS only)
    const std::string norm_file = disk + "_max_base_norm.bin";
    if (metric == 1 && exists(norm_file)) {
        File f(norm_file, err);
        if (!f.ok()) return false;
        std::vector<float> v;
        if (!f.bin<float>(0, &v, &r, &c, 16) || v.empty()) return false;
        o->max_base_norm = v[0];
    }
    return true;
}

}  // namespace lb2
