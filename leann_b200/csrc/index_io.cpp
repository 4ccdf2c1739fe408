// Reader for the reference's compact-CSR HNSW index file.
//
// Format = what leann_backend_hnsw/convert_to_csr.py:182-237 (write_compact_format) emits and
// the faiss fork's read_index reads (faiss/impl/index_read.cpp:1402-1490 "IHNf" header +
// read_HNSW :523-813):  header, assign_probas, cum_nneighbor_per_level, levels, compact flag,
// compact_level_ptr, compact_node_offsets, 5 x int32 scalars, storage fourcc,
// compact_neighbors_data, optional IndexFlat blob (index_write.cpp:419-426).
// Every offset is validated so that the GPU kernels can index without bounds checks.
#include "index_io.h"

#include <stdio.h>
#include <string.h>

namespace lb2 {

namespace {

constexpr uint32_t fourcc(const char (&s)[5]) {
    return (uint32_t)(uint8_t)s[0] | ((uint32_t)(uint8_t)s[1] << 8) | ((uint32_t)(uint8_t)s[2] << 16) |
           ((uint32_t)(uint8_t)s[3] << 24);
}

struct Reader {
    FILE* f;
    std::string* err;
    bool ok = true;
    bool raw(void* dst, size_t n) {
        if (!ok) return false;
        if (n && fread(dst, 1, n, f) != n) {
            *err = "unexpected end of file";
            ok = false;
        }
        return ok;
    }
    template <class T>
    bool one(T* v) { return raw(v, sizeof(T)); }
    template <class T>
    bool vec(std::vector<T>* v, uint64_t max_count) {
        uint64_t n = 0;
        if (!one(&n)) return false;
        if (n > max_count) {
            *err = "vector length " + std::to_string(n) + " exceeds sanity bound " + std::to_string(max_count);
            return ok = false;
        }
        {   // a count larger than what is left of the file is a corrupt header, not an allocation request
            const long here = ftell(f);
            if (here >= 0 && fseek(f, 0, SEEK_END) == 0) {
                const long end = ftell(f);
                fseek(f, here, SEEK_SET);
                if (end >= here && n * sizeof(T) > static_cast<uint64_t>(end - here)) {
                    *err = "unexpected end of file: a vector of " + std::to_string(n) + " elements does not fit in what is left";
                    return ok = false;
                }
            }
        }
        v->resize(n);
        return raw(v->data(), n * sizeof(T));
    }
};

}  // namespace

bool read_compact_index(const char* path, HostIndex* out, std::string* err) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        *err = std::string("cannot open index file '") + path + "'";
        return false;
    }
    Reader r{f, err};
    HostIndex& h = *out;
    uint32_t h4 = 0;
    int64_t dummy = 0;
    uint8_t is_trained = 0, compact = 0;
    r.one(&h4);
    if (r.ok && h4 != fourcc("IHNf")) {
        *err = "not an IndexHNSWFlat ('IHNf') file";
        r.ok = false;
    }
    r.one(&h.d);
    r.one(&h.ntotal);
    r.one(&dummy);
    r.one(&dummy);
    r.one(&is_trained);
    r.one(&h.metric_type);
    if (r.ok && h.metric_type > 1) r.one(&h.metric_arg);
    if (r.ok && (h.d <= 0 || h.ntotal < 0 || h.ntotal > 0x7fffffffLL)) {
        *err = "bad header (d / ntotal)";
        r.ok = false;
    }
    if (r.ok && h.metric_type != 0 && h.metric_type != 1) {
        *err = "unsupported metric_type " + std::to_string(h.metric_type) + " (only INNER_PRODUCT and L2)";
        r.ok = false;
    }
    const uint64_t N = r.ok ? (uint64_t)h.ntotal : 0;
    r.vec(&h.assign_probas, 1024);
    r.vec(&h.cum_nneighbor_per_level, 1024);
    r.vec(&h.levels, N);
    if (r.ok && h.levels.size() != N) {
        *err = "levels size != ntotal";
        r.ok = false;
    }
    r.one(&compact);
    if (r.ok && compact != 1) {
        *err = "index is not in compact CSR form (storage_is_compact flag missing)";
        r.ok = false;
    }
    r.vec(&h.level_ptr, N * 64 + 64);
    r.vec(&h.node_offsets, N + 1);
    int32_t scal[5] = {0, 0, 0, 0, 0};
    r.raw(scal, sizeof(scal));
    h.entry_point = scal[0];
    h.max_level = scal[1];
    h.ef_construction = scal[2];
    h.ef_search = scal[3];
    r.one(&h.storage_fourcc);
    r.vec(&h.neighbors, (uint64_t)1 << 40);
    if (r.ok && (h.storage_fourcc == fourcc("IxFI") || h.storage_fourcc == fourcc("IxF2"))) {
        int32_t d2 = 0, mt = 0;
        int64_t n2 = 0;
        uint8_t tr = 0;
        r.one(&d2); r.one(&n2); r.one(&dummy); r.one(&dummy); r.one(&tr); r.one(&mt);
        uint64_t nwords = 0;
        r.one(&nwords);
        if (r.ok && (d2 != h.d || n2 != h.ntotal || nwords != N * (uint64_t)h.d)) {
            *err = "flat storage blob does not match the index header";
            r.ok = false;
        }
        if (r.ok) {
            h.vectors.resize(nwords);
            r.raw(h.vectors.data(), nwords * 4);
        }
    } else if (r.ok && h.storage_fourcc != fourcc("null")) {
        *err = "unsupported storage index type (only 'null' / IndexFlat)";
        r.ok = false;
    }
    fclose(f);
    if (!r.ok) return false;

    // ---- structural validation
    if (h.node_offsets.size() != N + 1) { *err = "compact_node_offsets size != ntotal + 1"; return false; }
    if (N == 0) return true;
    if (h.node_offsets[N] != h.level_ptr.size()) { *err = "compact_node_offsets[ntotal] != |compact_level_ptr|"; return false; }
    if (h.entry_point < -1 || h.entry_point >= (int64_t)N) { *err = "entry_point out of range"; return false; }
    uint64_t prev = 0;
    for (uint64_t i = 0; i < N; i++) {
        const uint64_t ps = h.node_offsets[i], pe = h.node_offsets[i + 1];
        if (pe < ps || pe > h.level_ptr.size()) { *err = "compact_node_offsets not monotone"; return false; }
        for (uint64_t p = ps; p < pe; p++) {
            if (h.level_ptr[p] < prev || h.level_ptr[p] > h.neighbors.size()) { *err = "compact_level_ptr not monotone / out of range"; return false; }
            prev = h.level_ptr[p];
        }
    }
    for (int32_t v : h.neighbors)
        if (v < 0 || (uint64_t)v >= N) { *err = "neighbour id out of range"; return false; }
    if (h.entry_point >= 0) {
        const int nlev = (int)(h.node_offsets[h.entry_point + 1] - h.node_offsets[h.entry_point]) - 1;
        if (h.max_level < 0 || h.max_level >= (nlev > 0 ? nlev : 1)) { *err = "max_level inconsistent with the entry point's levels"; return false; }
    }
    return true;
}

}  // namespace lb2
