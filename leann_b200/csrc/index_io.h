// Host-side loader for the reference's compact-CSR HNSW .index file.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace lb2 {

struct HostIndex {
    int d = 0;
    int64_t ntotal = 0;
    int metric_type = 0;  // faiss MetricType: 0 = INNER_PRODUCT, 1 = L2
    float metric_arg = 0.f;
    std::vector<double> assign_probas;
    std::vector<int32_t> cum_nneighbor_per_level;
    std::vector<int32_t> levels;
    std::vector<uint64_t> level_ptr;
    std::vector<uint64_t> node_offsets;
    std::vector<int32_t> neighbors;
    int entry_point = -1, max_level = -1, ef_construction = 0, ef_search = 0;
    uint32_t storage_fourcc = 0;
    std::vector<float> vectors;  // filled when the file carries an IndexFlat storage blob
};

// Returns false and fills `err` on malformed input.
bool read_compact_index(const char* path, HostIndex* out, std::string* err);

}  // namespace lb2
