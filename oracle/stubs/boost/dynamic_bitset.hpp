/* TEST INFRASTRUCTURE (oracle/stubs) — stand-in for <boost/dynamic_bitset.hpp> (Boost is not installed here).
 * DiskANN's scratch.cpp only constructs, reset()s and deletes one for the IN-MEMORY index scratch, which the disk search
 * path under test (PQFlashIndex::cached_beam_search) never touches. */
#ifndef LB2_STUB_BOOST_DYNAMIC_BITSET_HPP
#define LB2_STUB_BOOST_DYNAMIC_BITSET_HPP
#include <cstddef>
#include <memory>
#include <vector>
namespace boost {
template <typename Block, typename Allocator>
class dynamic_bitset {
  public:
    dynamic_bitset() = default;
    void resize(std::size_t n) { bits_.assign(n, false); }
    void reset() { bits_.assign(bits_.size(), false); }
    std::size_t size() const { return bits_.size(); }
    bool operator[](std::size_t i) const { return bits_[i]; }
    std::vector<bool>::reference operator[](std::size_t i) { return bits_[i]; }

  private:
    std::vector<bool> bits_;
};
}  // namespace boost
#endif
