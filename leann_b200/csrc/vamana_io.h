// Host-side loader for the file set of the reference's DiskANN backend.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace lb2 {

struct VamanaHost {
    int64_t n = 0;
    int data_dim = 0;   // PQ ndims == stored coordinate count (raw dim + 1 for MIPS)
    int n_chunks = 0;
    int R = 0;          // max degree
    int metric = 0;     // 0 L2, 1 inner product, 2 cosine (diskann::Metric order used by the python binding)
    bool partitioned = false;
    std::vector<int32_t> nbrs;           // [n, R] valid ids first, padded with -1
    std::vector<uint8_t> codes;          // [n, n_chunks]
    std::vector<float> tables_tr;        // [data_dim, 256]
    std::vector<float> centroid;         // [data_dim]
    std::vector<uint32_t> chunk_offsets; // [n_chunks + 1]
    std::vector<uint32_t> medoids;
    std::vector<float> centroid_data;    // [n_medoids, data_dim] or empty
    std::vector<float> coords;           // [n, data_dim] or empty (partition mode never reads them)
    float max_base_norm = 0.f;
    int64_t n_edges = 0;
};

// product-quantisation tables + codes of a DiskANN-format file pair
struct PqHost {
    int64_t n = 0;
    int ndims = 0, n_chunks = 0;
    std::vector<uint8_t> codes;          // [n, n_chunks]
    std::vector<float> tables_tr;        // [ndims, 256]
    std::vector<float> centroid;         // [ndims]
    std::vector<uint32_t> chunk_offsets; // [n_chunks + 1]
};
bool read_pq_files(const std::string& pivots_path, const std::string& compressed_path, PqHost* out, std::string* err);

bool read_diskann_index(const char* index_prefix, const char* partition_prefix, int metric, VamanaHost* out, std::string* err);

}  // namespace lb2
