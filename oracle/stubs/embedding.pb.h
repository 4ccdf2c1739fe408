/* TEST INFRASTRUCTURE (oracle/stubs) — stand-in for the protoc-generated embedding.pb.h (no protoc / libprotobuf here).
 * The two messages of third_party/embedding.proto with exactly the members pq_flash_index.cpp:1583-1674 calls, encoded
 * in the proto3 wire format (packed repeated varints, length-delimited bytes), so the request the compiled reference
 * emits and the response it parses are byte-compatible with a real embedding server. */
#ifndef LB2_STUB_EMBEDDING_PB_H
#define LB2_STUB_EMBEDDING_PB_H
#include <cstdint>
#include <map>     // the protoc-generated header pulls these in transitively; pq_flash_index.cpp relies on it
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace protoembedding {
namespace wire {
inline void put_varint(std::string* s, uint64_t v) {
    while (v >= 0x80) { s->push_back(static_cast<char>((v & 0x7f) | 0x80)); v >>= 7; }
    s->push_back(static_cast<char>(v));
}
inline bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
    *v = 0;
    for (int shift = 0; p < end && shift < 64; shift += 7) {
        const uint8_t b = *p++;
        *v |= static_cast<uint64_t>(b & 0x7f) << shift;
        if (!(b & 0x80)) return true;
    }
    return false;
}
}  // namespace wire

class NodeEmbeddingRequest {
  public:
    void add_node_ids(uint32_t id) { node_ids_.push_back(id); }
    int node_ids_size() const { return static_cast<int>(node_ids_.size()); }
    uint32_t node_ids(int i) const { return node_ids_[i]; }
    bool SerializeToString(std::string* out) const {
        out->clear();
        if (node_ids_.empty()) return true;
        std::string payload;
        for (uint32_t id : node_ids_) wire::put_varint(&payload, id);
        out->push_back(static_cast<char>((1 << 3) | 2));  // field 1, length-delimited (packed)
        wire::put_varint(out, payload.size());
        out->append(payload);
        return true;
    }
    bool ParseFromArray(const void* data, int size) {
        node_ids_.clear();
        const uint8_t* p = static_cast<const uint8_t*>(data);
        const uint8_t* end = p + size;
        while (p < end) {
            uint64_t key, v;
            if (!wire::get_varint(p, end, &key)) return false;
            if (key == ((1 << 3) | 2)) {
                uint64_t len;
                if (!wire::get_varint(p, end, &len) || len > static_cast<uint64_t>(end - p)) return false;
                const uint8_t* pe = p + len;
                while (p < pe) { if (!wire::get_varint(p, pe, &v)) return false; node_ids_.push_back(static_cast<uint32_t>(v)); }
            } else if (key == ((1 << 3) | 0)) {
                if (!wire::get_varint(p, end, &v)) return false;
                node_ids_.push_back(static_cast<uint32_t>(v));
            } else {
                return false;
            }
        }
        return true;
    }

  private:
    std::vector<uint32_t> node_ids_;
};

class NodeEmbeddingResponse {
  public:
    const std::string& embeddings_data() const { return data_; }
    void set_embeddings_data(const void* p, size_t n) { data_.assign(static_cast<const char*>(p), n); }
    int dimensions_size() const { return static_cast<int>(dims_.size()); }
    int32_t dimensions(int i) const { return dims_[i]; }
    void add_dimensions(int32_t v) { dims_.push_back(v); }
    int missing_ids_size() const { return static_cast<int>(missing_.size()); }
    bool SerializeToString(std::string* out) const {
        out->clear();
        if (!data_.empty()) {
            out->push_back(static_cast<char>((1 << 3) | 2));
            wire::put_varint(out, data_.size());
            out->append(data_);
        }
        if (!dims_.empty()) {
            std::string payload;
            for (int32_t v : dims_) wire::put_varint(&payload, static_cast<uint64_t>(static_cast<int64_t>(v)));
            out->push_back(static_cast<char>((2 << 3) | 2));
            wire::put_varint(out, payload.size());
            out->append(payload);
        }
        return true;
    }
    bool ParseFromArray(const void* data, int size) {
        data_.clear(); dims_.clear(); missing_.clear();
        const uint8_t* p = static_cast<const uint8_t*>(data);
        const uint8_t* end = p + size;
        while (p < end) {
            uint64_t key, v, len;
            if (!wire::get_varint(p, end, &key)) return false;
            const int field = static_cast<int>(key >> 3), wt = static_cast<int>(key & 7);
            if (wt == 2) {
                if (!wire::get_varint(p, end, &len) || len > static_cast<uint64_t>(end - p)) return false;
                const uint8_t* pe = p + len;
                if (field == 1) data_.assign(reinterpret_cast<const char*>(p), len);
                else if (field == 2) { const uint8_t* q = p; while (q < pe) { if (!wire::get_varint(q, pe, &v)) return false; dims_.push_back(static_cast<int32_t>(v)); } }
                else if (field == 3) { const uint8_t* q = p; while (q < pe) { if (!wire::get_varint(q, pe, &v)) return false; missing_.push_back(static_cast<uint32_t>(v)); } }
                p = pe;
            } else if (wt == 0) {
                if (!wire::get_varint(p, end, &v)) return false;
                if (field == 2) dims_.push_back(static_cast<int32_t>(v));
                else if (field == 3) missing_.push_back(static_cast<uint32_t>(v));
            } else {
                return false;
            }
        }
        return true;
    }

  private:
    std::string data_;
    std::vector<int32_t> dims_;
    std::vector<uint32_t> missing_;
};
}  // namespace protoembedding
#endif
